"""B200-native batched Leduc Hold'em + NFSP rollout engine.

Host side of libnfsp_b200.so (hand-written sm_100a CUDA behind the C ABI in include/nfsp_b200.h).
The directory name is not an importable identifier; load it through the `nfsp_b200` alias module at
the repository root (`import nfsp_b200`), or call `install_dropin()` to expose the reference's own
module names (`leduc.env`, `leduc.newenv`, `agent.agent`, `utils.replay_buffer`,
`utils.ReservoirBuffer`) so that `import leduc.newenv as leduc` (main.py:3) resolves here.
"""
import sys as _sys

from ._lib import LIB_PATH, NfspError, SYMBOLS, lib  # noqa: F401
from .batched import (BatchedLegacyEnv, BatchedNfspEnv, DeviceReservoir, DeviceRing, SelfPlay,  # noqa: F401
                      decode_trace, expand_obs, glorot_nets, obs_to_mask, split_net)
from .config import Config, load_config  # noqa: F401

_DROPIN = ("leduc", "leduc.env", "leduc.newenv", "leduc.deck", "leduc.cardmatrix", "agent", "agent.agent", "utils",
           "utils.replay_buffer", "utils.ReservoirBuffer")


def install_dropin():
    """Register this package's reference-named modules at top level (see module docstring)."""
    import importlib

    for name in _DROPIN:
        _sys.modules[name] = importlib.import_module(__name__ + "." + name)
