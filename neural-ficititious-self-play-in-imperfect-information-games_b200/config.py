"""config.ini loader with the reference's keys and defaults (config.ini:1-40).

The reference re-reads "./config.ini" from the CWD in four places and even inside hot calls
(env.py:58,119,135); here it is read once on the host and the values travel as kernel arguments.
"""
from __future__ import annotations

import configparser
import os

DEFAULTS = {
    "Agent": dict(HiddenLayer=64, LearningRateBR=0.05, LearningRateAR=0.1, Gamma=0.95, Epsilon=0.06,
                  EpsilonDecay=0.06, EpsilonMin=0, MiniBatchSize=128, Penalty=-1, Eta=0.1, MRLSize=200000,
                  MSLSize=2000000, Omega=0.003, TargetModelUpdateRate=150),
    "Environment": dict(Decksize=6, Playercount=2, Choices=4, MaxRounds=2, Suits=2, MaxRaises=3, ActionSpace=2,
                        TotalActionSpace=3),
    "Utils": dict(Buffersize=40000, Seed=1234),
    "Common": dict(MaxEpisodes=10000, Episodes=400000, TestEpisodes=1000),
}


class Config:
    """`ConfigParser.ConfigParser`-shaped accessor: get(section, key) returns a string."""

    def __init__(self, path: str | None = "./config.ini"):
        self._c = configparser.ConfigParser()
        self._c.optionxform = str
        for sec, kv in DEFAULTS.items():
            self._c[sec] = {k: str(v) for k, v in kv.items()}
        if path and os.path.isfile(path):
            self._c.read(path)

    def read(self, path):
        if os.path.isfile(path):
            self._c.read(path)

    def get(self, section, key):
        return self._c.get(section, key)

    def getint(self, section, key):
        return int(float(self._c.get(section, key)))

    def getfloat(self, section, key):
        return float(self._c.get(section, key))


def load_config(path: str | None = "./config.ini") -> Config:
    cfg = Config(path)
    env = cfg
    # the kernels hard-wire the game the reference ships (config.ini:19-29); refuse anything else loudly
    fixed = dict(Decksize=6, Playercount=2, Choices=4, MaxRounds=2, Suits=2, MaxRaises=3, ActionSpace=2,
                 TotalActionSpace=3)
    for k, v in fixed.items():
        if env.getint("Environment", k) != v:
            raise ValueError("Environment.%s=%s is not supported by the sm_100a kernels (needs %d)" %
                             (k, env.get("Environment", k), v))
    if cfg.getint("Agent", "Penalty") != -1 or cfg.getint("Agent", "HiddenLayer") != 64:
        raise ValueError("Agent.Penalty must be -1 and Agent.HiddenLayer 64 for the sm_100a kernels")
    return cfg
