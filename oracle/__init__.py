"""CPU oracle: a restatement of the reference's algorithm for the LPG meta-training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``to_ued_b200/`` may import this package.  The only
callers are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and there only as the checker / the timed CPU baseline.

PARITY UNPINNED.  The reference (nmonette/TO-UED, a JAX program) has no tests, golden vectors or
fixtures, and jax / flax / optax / evosax / gymnax are not installable in this
image, so the reference cannot be run here.  This oracle follows the reference files line by line
(each function cites ``file:line``) and restates the published algorithms of the un-vendored
third-party pieces it needs:

  * jax==0.4.13  threefry2x32 PRNG, ``split`` / ``uniform`` / ``bernoulli`` / ``choice`` /
    ``permutation`` lowering                                         (oracle/prng.py)
  * gymnax==0.0.6 ``Environment.step`` key split + auto-reset          (oracle/gridworld.py)
  * flax==0.6.11 ``Dense`` / ``GRUCell`` gate layout and initialisers   (oracle/lpg.py)
  * optax==0.1.5 ``clip_by_global_norm`` / ``scale_by_adam``            (oracle/optim.py)
  * evosax==0.1.4 ``OpenES`` ask / tell                                 (oracle/es.py)

What *is* pinned: the threefry2x32 block function against the three Random123 known-answer
vectors (the same ones jax's own test-suite uses) and the ``split(PRNGKey(0))`` /
``uniform(PRNGKey(0))`` values printed in the public JAX documentation
(tests/test_01_oracle_prng.py).  Everything downstream of those is "[3P-recall]".

Layout
------
``prng.py``        counter-based RNG contract shared bit-for-bit with the CUDA kernels
``gridworld.py``   EnvParams / EnvState, step_env / reset_env / get_obs   (integer, bit-exact)
``rollout.py``     batch_reset / batch_rollout / single_rollout          (integer, bit-exact)
``configs.py``     level generator + per-mode tables
``lpg.py``         LPG network (embed MLP, reverse GRU, heads) in torch (CPU, autograd)
``agents.py``      actor / critic tables, LPG-driven and A2C updates, eval_agent, GAE
``meta.py``        lpg_meta_grad_train_step (torch.autograd through the K updates), ES step
``level_sampler.py`` PLR buffer logic (bit-exact index work)
"""
