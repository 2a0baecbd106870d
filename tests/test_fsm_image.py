"""CPU check of the table image behind the state-machine step kernel (csrc/nfsp_fsm.cuh, nfsp_fsm_image).

All of newenv.Env.step (newenv.py:131-349) and main.train's turn order (main.py:28-67) live in that image, so it is
walked here on the host with the kernel's own per-step recipe -- the two table loads, the masks, the swap -- and the
resulting 12-byte trace records are held against the CPU restatement of the reference on the same Philox streams.
Tolerance: none, every word is compared bit for bit (the reward plane as float bits)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural-ficititious-self-play-in-imperfect-information-games_b200")

# layout constants of csrc/nfsp_fsm.cuh
ROW, ENTRY, INFO = 112, 32, 96
LIVE_ROWS, ROWS = 36, 52
DEAL_OFF = ROWS * ROW
POL_OFF = 120 * 16
DEAL_HALF = POL_OFF + 4 * 16
REWARD_OFF = DEAL_OFF + 2 * DEAL_HALF
P_CLEAR = 0x7E000
CM0 = 0x07000000
OVER = 1 << 30
M32 = 0xFFFFFFFF


@pytest.fixture(scope="module")
def image():
    import __graft_entry__ as g

    if not os.path.exists(os.path.join(PKG, "libnfsp_b200.so")):
        g.build()
    import nfsp_b200

    L = nfsp_b200.lib()
    n = L.nfsp_fsm_image(None, 0)
    assert n * 4 == REWARD_OFF + 4 * 1024
    buf = np.zeros(n, np.uint32)
    assert L.nfsp_fsm_image(buf.ctypes.data_as(ctypes.c_void_p), n) == n
    return buf


def words(img, byte_off, count):
    return [int(v) for v in img[byte_off // 4: byte_off // 4 + count]]


class Walker:
    """One game, stepped the way nfsp_step_fsm_kernel steps it (addresses are byte offsets into the image)."""

    def __init__(self, img, seed, game, eta_u32):
        self.img, self.game, self.eta = img, game, eta_u32
        self.key = (seed & M32, seed >> 32)

    def block(self, step):
        return [int(v) for v in orc.philox((self.game & M32, step & M32, step >> 32, 0), self.key)]

    def deal(self, x, dealer):
        self.dl = DEAL_OFF + dealer * DEAL_HALF
        D = words(self.img, self.dl + ((x[1] * 120) >> 32) * 16, 4)
        pa = self.dl + (16 if x[2] < self.eta else 0) + (32 if x[3] < self.eta else 0)
        P = words(self.img, pa + POL_OFF, 4)
        self.PA, self.PO, self.ms, self.tix, self.HX = D[0] | P[0], D[1] | P[1], D[2] | P[2], P[3], CM0

    def reset(self, step):  # nfsp_reset_kernel: game g starts with dealer g & 1
        self.deal(self.block(step), self.game & 1)

    def step(self, step):
        x = self.block(step)
        started = 0
        if self.HX & OVER:
            self.deal(x, 1 - (self.dl - DEAL_OFF) // DEAL_HALF)
            started = 1 << 21
        ea = self.tix + ((x[0] * 3) >> 32) * ENTRY
        hc, pa, pm, nx, mw, mb, ma, mo = words(self.img, ea, 8)
        self.PA = ((self.PA & ~P_CLEAR) + pa) & M32
        self.HX = (self.HX & 0xFFFFFF) | hc
        self.ms = (self.ms + mb) & M32
        obs = self.HX & (self.PA | 0xC0FFFFFF)
        reward = words(self.img, REWARD_OFF + ((self.PA & ma) | (self.PO & mo)), 1)[0]
        misc = self.ms | mw | started
        a = self.PA
        self.PA = (self.PO & pm) | (a & ~pm & M32)
        self.PO = (a & pm) | (self.PO & ~pm & M32)
        self.tix = nx
        return obs, reward, misc


@pytest.mark.parametrize("seed,eta", [(1234, 0.1), (0xDEADBEEFCAFE, 0.5)])
def test_walking_the_image_reproduces_the_oracle_traces(image, seed, eta):
    n, steps = 160, 48
    eta_u32 = orc.u32_frac(eta)
    b = orc.NfspBatch(n, seed)
    b.reset(0, eta_u32)
    ref = b.rollout_env(1, steps, eta_u32)
    hands = 0
    for g in range(n):
        w = Walker(image, seed, g, eta_u32)
        w.reset(0)
        for t in range(steps):
            obs, reward, misc = w.step(1 + t)
            where = "game %d step %d" % (g, t)
            assert obs == int(ref["obs"][t, g]), where
            assert reward == int(ref["reward"][t, g].view(np.uint32)), where
            assert misc == int(ref["misc"][t, g]), where
            hands += (obs >> 30) & 1
    assert hands > n * steps // 4  # about 2.6 transitions per hand: the re-deal path was walked many times


def test_image_structure(image):
    """Every entry of a row that can be acted in points at a row of the image; a terminating entry points at a
    terminal pseudo-row whose info word says so; the turn passes exactly when the next row's actor differs."""
    seen_term = set()
    for row in range(LIVE_ROWS):
        sigma = row % 18
        info = int(image[(row * ROW + INFO) // 4])
        if sigma % 9 >= 7:
            assert not image[row * ROW // 4: (row + 1) * ROW // 4].any()
            continue
        q = (info >> 4) & 1
        for raw in range(3):
            hc, pa, pm, nx, mw, mb, ma, mo = words(image, row * ROW + raw * ENTRY, 8)
            assert nx % ROW == 0 and nx // ROW < ROWS
            over = (hc >> 30) & 1
            assert (hc >> 31) == q and (mw & 3) == raw and ((pa >> 13) & 3) == raw
            nxt = int(image[(nx + INFO) // 4])
            assert ((nxt >> 16) & 1) == over
            if over:
                assert nx // ROW >= LIVE_ROWS and pm == 0 and ((nxt >> 4) & 1) == q
                seen_term.add(nx // ROW)
                assert (ma, mo) in ((0x3C, 0), (0xC0, 0xF00))
            else:
                assert nx // ROW < LIVE_ROWS and (nx // ROW) // 18 == row // 18  # the dealer does not change mid-hand
                assert pm == (M32 if ((nxt >> 4) & 1) != q else 0)
                assert ma == 0 and mo == 0
    assert len(seen_term) == 16
    rewards = image[REWARD_OFF // 4:].view(np.float32)
    assert rewards[0] == 0.0 and not np.signbit(rewards[0])
    assert set(np.unique(np.abs(rewards))) <= {0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 4.5, 5.0, 5.5, 6.0, 6.5, 7.0, 7.5}


# ------------------------------------------------------------------------------------------- legacy rules
L_ENTRY, L_ROW, L_MAXROWS = 48, 9 * 48, 72
L_KEY_OFF = L_MAXROWS * L_ROW
L_DEAL_OFF = L_KEY_OFF + 4096
L_TERM = 1 << 24


@pytest.fixture(scope="module")
def legacy_image(image):
    import nfsp_b200

    L = nfsp_b200.lib()
    n = L.nfsp_legacy_fsm_image(None, 0, None)
    assert n * 4 == L_DEAL_OFF + 120 * 16
    buf = np.zeros(n, np.uint32)
    rows = ctypes.c_int(0)
    assert L.nfsp_legacy_fsm_image(buf.ctypes.data_as(ctypes.c_void_p), n, ctypes.byref(rows)) == n
    return buf, rows.value


def s32(v):
    return v - (1 << 32) if v & 0x80000000 else v


def test_walking_the_legacy_image_reproduces_the_oracle_records(legacy_image):
    """legacy_rollout_kernel's per-iteration recipe on the host against the restated leduc/env.py README loop."""
    img, rows = legacy_image
    assert rows == 70  # the states reachable from a fresh hand (left, pots, carried penalties)
    n, iters, seed = 200, 40, 4321
    b = orc.LegacyBatch(n, seed)
    b.reset(0)
    ref = b.rollout(1, iters)
    key = (seed & M32, seed >> 32)
    keys = img.view(np.uint8)[L_KEY_OFF: L_KEY_OFF + 4096]
    assert int((keys != 0xFF).sum()) == rows and keys[5 | 5 << 3] == 0  # a fresh hand (left 4/4, pots 0/0) is row 0
    live_keys = set()
    ended = 0
    for g in range(n):
        x = [int(v) for v in orc.philox((g, 0, 0, 0), key)]
        c0, c1, G, L = words(img, L_DEAL_OFF + ((x[2] * 120) >> 32) * 16, 4)
        six, need = 0, 0
        for t in range(iters):
            x = [int(v) for v in orc.philox((g, 1 + t, 0, 0), key)]
            started = 0
            if need:
                c0, c1, G, L = words(img, L_DEAL_OFF + ((x[2] * 120) >> 32) * 16, 4)
                six, started = 0, 1 << 8
            a0, a1 = (x[0] * 3) >> 32, (x[1] * 3) >> 32
            nx, w0, rb0, rb1, w2x, w2y, p0t, p1t, lo = words(img, six + (a0 * 3 + a1) * L_ENTRY, 9)
            u = L * p1t - G * p0t
            r = (s32(rb0) + u, s32(rb1) - u)
            for p, (c, w2) in enumerate(((c0, w2x), (c1, w2y))):
                rec, where = ref[t, g, p], "game %d iteration %d player %d" % (g, t, p)
                assert (c, -1, (w0 >> 16) & 0xFF, (w0 >> 24) & 1) == (rec["card"], rec["pub"], rec["pot"], rec["terminal"]), where
                assert r[p] == rec["reward"] and (w2 | started) == int(rec["misc"]), where
            need = w0 & L_TERM
            ended += 1 if need else 0
            assert need or nx % L_ROW == 0 and nx // L_ROW < rows
            if not need:
                # the state-key table (packed word -> row, used when a live hand crosses a launch boundary): the key of
                # the state the ORACLE is in now must lead to the row the entry points at
                skey = 0
                for p in (0, 1):
                    m = int(ref[t, g, p]["misc"])
                    skey |= ((m >> 2) & 7) << (3 * p) | ((m >> 5) & 7) << (6 + 2 * p) | (1 if ref[t, g, p]["reward"] != 0 else 0) << (10 + p)
                assert int(keys[skey]) == nx // L_ROW, "game %d iteration %d" % (g, t)
                live_keys.add(skey)
            six = nx
    assert ended > n * iters // 8 and len(live_keys) > 20
