"""Drop-in counterparts of the reference's ``lib`` package (only the hot-path modules)."""
