"""Run the UNMODIFIED reference sources from /root/reference under Python 3 / numpy 2.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
use it.  This module additionally needs /root/reference, which exists only in the
build container -- it is used solely by oracle/gen_golden.py to produce the
committed fixtures under tests/golden/ (and by tests that skip when it is absent).

The reference is Python 2.7.  No reference source is copied
into this repo: the files are read from REF_ROOT at run time, five one-token
Py2->Py3 / numpy-2 edits are applied to the text IN MEMORY, and the result is
exec'd as modules named like the originals:

  leduc/deck.py:37        `self._size / 2`            -> `//`   (Py2 int division)
  leduc/cardmatrix.py:7-8 ragged np.array             -> dtype=object
  leduc/env.py:64-70      ragged np.array             -> dtype=object
  leduc/newenv.py:39,106  `self.decksize / self.suits`-> `//`   (Py2 int division)

plus a `ConfigParser` shim module (Py2 name) over `configparser`.
`agent/agent.py` and `main.py` import keras / tensorflow, which are not installed
and whose arithmetic is third-party; their *control flow* (Agent.play & friends,
main.train) is extracted with `ast` and exec'd against stub models so that the
turn order, RL/SL memory gating and buffer traffic come from the reference's own
code, not from a restatement.
"""
from __future__ import annotations

import ast
import configparser
import os
import random
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("NFSP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "leduc", "newenv.py"))


def _read(rel: str) -> str:
    with open(os.path.join(REF_ROOT, rel), "r") as f:
        return f.read()


def _replace_once(text: str, old: str, new: str, count: int = 1) -> str:
    assert text.count(old) == count, (old, text.count(old))
    return text.replace(old, new)


class _Cfg:
    """`ConfigParser.ConfigParser` stand-in reading the reference's own config.ini
    from REF_ROOT regardless of CWD (reference reads "./config.ini")."""

    def __init__(self):
        self._c = configparser.ConfigParser()

    def read(self, path):
        self._c.read(os.path.join(REF_ROOT, "config.ini"))

    def get(self, section, key):
        return self._c.get(section, key)


_MODS: dict | None = None


def load():
    """Returns dict of patched reference modules: cardmatrix, deck, env, newenv,
    replay_buffer, ReservoirBuffer."""
    global _MODS
    if _MODS is not None:
        return _MODS
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)

    shim = types.ModuleType("ConfigParser")
    shim.ConfigParser = _Cfg
    sys.modules["ConfigParser"] = shim

    srcs = {}
    t = _read("leduc/cardmatrix.py")
    t = _replace_once(t, "['Heart', 'Spades', 'Cross', 'Diamonds']])",
                      "['Heart', 'Spades', 'Cross', 'Diamonds']], dtype=object)")
    srcs["cardmatrix"] = t
    t = _read("leduc/deck.py")
    t = _replace_once(t, "self._size / 2", "self._size // 2")
    srcs["deck"] = t
    t = _read("leduc/env.py")
    t = _replace_once(t, "self._info                                  # info\n            ])",
                      "self._info                                  # info\n            ], dtype=object)")
    srcs["env"] = t
    t = _read("leduc/newenv.py")
    t = _replace_once(t, "(self.decksize / self.suits)", "(self.decksize // self.suits)", 2)
    srcs["newenv"] = t
    srcs["replay_buffer"] = _read("utils/replay_buffer.py")
    srcs["ReservoirBuffer"] = _read("utils/ReservoirBuffer.py")

    mods = {}
    saved = {k: sys.modules.get(k) for k in ("cardmatrix", "deck")}
    for name in ("cardmatrix", "deck", "env", "newenv", "replay_buffer", "ReservoirBuffer"):
        m = types.ModuleType("_nfsp_ref_" + name)
        m.__file__ = os.path.join(REF_ROOT, name + ".py")
        # the reference uses Py2 implicit-relative `import deck` / `import cardmatrix`
        if name in ("cardmatrix", "deck"):
            sys.modules[name] = m
        exec(compile(srcs[name], m.__file__, "exec"), m.__dict__)
        mods[name] = m
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    _MODS = mods
    return mods


# ---------------------------------------------------------------------------
# deck injection: cards are popped from the END of `_cards`
# ---------------------------------------------------------------------------
def inject_deck(deck_obj, ranks_in_pop_order):
    """Make `deck_obj.pick_up()` yield the given ranks in order (p0, p1, public...)."""
    mods = load()
    Card = mods["deck"].Card
    cards = [Card(int(r), 0) for r in ranks_in_pop_order]
    deck_obj._cards = list(reversed(cards))


class ScriptedShuffle:
    """Context manager: while active, `Deck.shuffle()` deals a scripted rank order
    instead of calling random.shuffle (reference deck.py:42-44)."""

    def __init__(self, ranks_in_pop_order_fn):
        self.fn = ranks_in_pop_order_fn

    def __enter__(self):
        mods = load()
        self.Deck = mods["deck"].Deck
        self.orig = self.Deck.shuffle
        fn = self.fn

        def shuffle(deck_self):
            inject_deck(deck_self, fn())

        self.Deck.shuffle = shuffle
        return self

    def __exit__(self, *a):
        self.Deck.shuffle = self.orig


# ---------------------------------------------------------------------------
# control flow of agent.Agent / main.train, extracted from the reference text
# ---------------------------------------------------------------------------
_AGENT_METHODS = ("remember_best_response", "remember_for_rl", "act_best_response", "play",
                  "boltzmann", "sampled_actions")


def _extract_functions(src: str, names, class_name=None):
    tree = ast.parse(src)
    out = []
    body = tree.body
    if class_name is not None:
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name == class_name:
                body = node.body
    for node in body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            out.append(node)
    assert len(out) == len(names), [n.name for n in out]
    return out


def agent_class(random_module=random, np_module=np):
    """A class whose play / act_best_response / remember_* / boltzmann methods are the
    reference's own (agent/agent.py:118-166), compiled from its source text; the ctor
    (which builds Keras models) is replaced by a stub taking explicit models+buffers."""
    src = _read("agent/agent.py")
    fns = _extract_functions(src, _AGENT_METHODS, class_name="Agent")
    mod = ast.Module(body=[ast.ClassDef(name="Agent", bases=[], keywords=[], body=fns, decorator_list=[],
                                        type_params=[])], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"np": np_module, "random": random_module, "print": lambda *a, **k: None}
    exec(compile(mod, os.path.join(REF_ROOT, "agent/agent.py"), "exec"), ns)
    Base = ns["Agent"]

    class StubAgent(Base):
        def __init__(self, name, env, avg_model, br_model, rl_memory, sl_memory, epsilon=0.06, temp=1.0,
                     on_update=None):
            self.name = name
            self.env = env
            self.avg_strategy_model = avg_model
            self.best_response_model = br_model
            self._rl_memory = rl_memory
            self._sl_memory = sl_memory
            self.epsilon = epsilon
            self.temp = temp
            self.actions = np.zeros(3)
            self.played = 0
            self.reward = 0
            self.game_step = 0
            self._on_update = on_update
            self.updates = 0

        def update_strategy(self):  # learner is out of the rollout path
            self.updates += 1
            if self._on_update is not None:
                self._on_update(self)

    return StubAgent


def train_function(episodes: int, eta: float, random_module=random):
    """main.train (main.py:21-124) compiled from the reference text with a Config stub
    (Episodes, Eta), no matplotlib and no sleep."""
    src = _read("main.py")
    (fn,) = _extract_functions(src, ("train",))
    mod = ast.Module(body=[fn], type_ignores=[])
    ast.fix_missing_locations(mod)

    class _Config:
        @staticmethod
        def get(section, key):
            if (section, key) == ("Agent", "Eta"):
                return str(eta)
            if (section, key) == ("Common", "Episodes"):
                return str(episodes)
            raise KeyError((section, key))

    class _Plt:
        @staticmethod
        def plot(*a, **k):
            pass

        @staticmethod
        def show(*a, **k):
            pass

    class _Time:
        @staticmethod
        def sleep(*a):
            pass

    ns = {"np": np, "random": random_module, "Config": _Config, "plt": _Plt, "time": _Time,
          "print": lambda *a, **k: None}
    exec(compile(mod, os.path.join(REF_ROOT, "main.py"), "exec"), ns)
    return ns["train"]
