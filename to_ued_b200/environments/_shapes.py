def obs_dim(max_grid_size: int, max_n_objs: int) -> int:
    """Tabular observation size, reference gridworld.py:230-233."""
    return max_grid_size ** 2 * (2 ** max_n_objs) + 1
