"""`modules.functional` alias, so `from <pkg>.modules import functional as F` resolves like in the
reference tree (modules/functional/ lives inside modules/ there)."""
from ...functional import *  # noqa: F401,F403
from ...functional import _backend  # noqa: F401
