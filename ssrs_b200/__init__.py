"""ssrs_b200 — B200-native implementation of the SSRS hot path behind the reference's Config/Simulator API.

    from ssrs_b200 import Config, Simulator
    sim = Simulator(Config(...), elevation=dem)     # terrain injected; the reference downloads it
    sim.simulate_tracks()

The compute path is hand-written sm_100a CUDA behind the C-ABI of `include/ssrs_b200.h`
(`ssrs_b200/libssrs_b200.so`, built by `python -m ssrs_b200.build`).  There is no CPU fallback.
"""
from .config import Config  # noqa: F401

__all__ = ["Config", "Simulator"]


def __getattr__(name):
    if name == "Simulator":
        from .simulator import Simulator
        return Simulator
    raise AttributeError(name)
