"""``lib.models`` call surface of the reference (networks / loss / util), backed by the sm_100a kernels."""
from . import loss, networks, util  # noqa: F401
