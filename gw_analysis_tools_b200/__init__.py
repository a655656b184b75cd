"""B200-native batched likelihood engine behind gw_analysis_tools' own entry points."""
from . import abi  # noqa: F401
