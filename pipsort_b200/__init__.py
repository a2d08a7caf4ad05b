"""pipsort_b200: B200-native (sm_100a) posterior-calculation engine behind PIPSORT's PostCal boundary."""
from .engine import Engine, PipsortError, Results, measure_fp64_peak, version, lib, KEEP_ORDER, KMAX  # noqa: F401
