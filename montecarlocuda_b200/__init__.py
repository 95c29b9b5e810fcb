"""montecarlocuda_b200 -- B200-native (sm_100a) Monte Carlo pricing behind the reference's C host API.

The product is the CUDA library in `lib/` (built from `csrc/`, ABI in `include/`); this package is
the thin host-side mirror of the reference interface used by the tests and the benchmark.
"""
from ._lib import F32, F64, Mcb200Error, library_path, load  # noqa: F401
from .api import (BASKET_FFMA, BASKET_TENSOR, get_basket_engine, set_basket_engine,  # noqa: F401
                  CVA, DEFAULT_SEED, Engine, MultiOptionData, OptionData, OptionValue,  # noqa: F401
                  dev_basketOpt, dev_cvaEquityOption, dev_vanillaOpt, finalize, plan, shard_range)

__all__ = ["BASKET_FFMA", "BASKET_TENSOR", "get_basket_engine", "set_basket_engine", "F32", "F64", "Mcb200Error", "library_path", "load", "CVA", "DEFAULT_SEED", "Engine",
           "MultiOptionData", "OptionData", "OptionValue", "dev_basketOpt", "dev_cvaEquityOption",
           "dev_vanillaOpt", "finalize", "plan", "shard_range"]
