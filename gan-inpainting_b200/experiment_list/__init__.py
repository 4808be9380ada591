"""``begin(state, loaders)`` recipes with the reference's call surface (experiment_list/*.py)."""
