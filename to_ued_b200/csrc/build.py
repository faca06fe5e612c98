"""Build libtoued.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and importable
as ``python -m to_ued_b200.csrc.build``.  One object per .cu so only changed files recompile."""
import os, subprocess, sys, hashlib, shutil

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "libtoued.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
# per-file overrides: the sampling path must round exactly like the oracle (no FMA contraction),
# so rollout.cu adds -fmad=false.
FLAGS = {
    "rollout.cu": ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-fmad=false"],
    "levelgen.cu": ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-fmad=false"],
}


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


# diagnostic variants: name -> extra -D flags; built into libtoued_<name>.so next to the production library
VARIANTS = {
    # delays the odd-pass epilogue warps of gru_forward_tc at pass 13 so that an x-tile write-after-read hazard
    # (if present) fires on every step (tests/diag_nondeterminism.py)
    "probe": ["-DTOUED_RACE_PROBE=1"],
    # random delays in front of every mbarrier wait: schedule fuzzing of the warp-specialised kernels
    "fuzz": ["-DTOUED_FUZZ=1"],
    "fuzz_old": ["-DTOUED_FUZZ=1", "-DTOUED_XTILE_OLD=1"],
    "bwd16": ["-DBT_EW=16"],                 # 16 epilogue warps in gru_backward_tc (measured slower, see the kernel's header)
    # weight-gradient GEMM experiments of round 2 (wgrad_tc.cu; both correct on every parity test, both measured slower):
    "wgpair": ["-DTOUED_WGRAD_PAIR=1"],      # cta_group::2: CTA pairs sharing M256 N256 MMAs
    "wpprobe": ["-DTOUED_WGRAD_PAIR=1", "-DTOUED_WP_PROBE=1"],   # the same with a launch-size override for timing experiments
    "wgk": ["-DTOUED_WGRAD_KMAJOR=1"],       # K-major operands transposed in shared memory (ldmatrix.trans / stmatrix)
    "fwdns7": ["-DFT_NS_OVERRIDE=7"],        # gru_forward_tc with 7 instead of 6 weight stages in flight
    "fwd1": ["-DFT_HEADS_WARP=0"],           # gru_forward_tc with the heads MMAs issued by the gate-MMA warp (round-1 arrangement)
    "pf2": ["-DBT_PF_DEPTH=2"],              # gru_backward_tc with two chunks of saved activations in flight per thread
    "pf": ["-DBT_L2_PREFETCH=1"],            # gru_backward_tc with an L2 bulk prefetch of the next step's saved activations (measured slower)
    "old": ["-DTOUED_XTILE_OLD=1"],
    "probe_old": ["-DTOUED_RACE_PROBE=1", "-DTOUED_XTILE_OLD=1"],      # the round-1 x-tile writer (set 0)
}


def build(verbose=False, force=False, variant=""):
    global OUT
    if variant:
        return _build(verbose, force, os.path.join(HERE, "..", f"libtoued_{variant}.so"),
                      os.path.join(HERE, f"build_{variant}"), VARIANTS[variant])
    return _build(verbose, force, OUT, os.path.join(HERE, "build"), [])


def _build(verbose, force, OUT, objdir, extra):
    srcs = sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))
    hdrs = sorted(f for f in os.listdir(HERE) if f.endswith(".cuh")) + ["../../include/toued.h"]
    hsig = hashlib.sha1(b"".join(open(os.path.join(HERE, h), "rb").read() for h in hdrs)).hexdigest()
    os.makedirs(objdir, exist_ok=True)
    objs, rebuilt = [], False
    for s in srcs:
        src = os.path.join(HERE, s)
        obj = os.path.join(objdir, s[:-3] + ".o")
        flags = list(FLAGS.get(s, COMMON)) + list(extra)
        sig = hashlib.sha1(open(src, "rb").read() + hsig.encode() + " ".join(flags).encode()).hexdigest()
        sigf = obj + ".sig"
        if force or not os.path.exists(obj) or not os.path.exists(sigf) or open(sigf).read() != sig:
            cmd = [_nvcc(), *ARCH, *flags, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            open(sigf, "w").write(sig)
            rebuilt = True
        objs.append(obj)
    if rebuilt or not os.path.exists(OUT):
        cmd = [_nvcc(), *ARCH, "-shared", "-o", OUT, *objs, "-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return os.path.abspath(OUT)


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, variant=var[0] if var else ""))
