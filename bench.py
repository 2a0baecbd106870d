#!/usr/bin/env python
"""Headline benchmark: Leduc transitions/sec of the fused NFSP rollout (env step + NFSP act + memories).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port on all host threads

One "step" = one pass of the hot path over one batch: `rollout(T)` (T decisions for every game: observe,
remember, act with the eta-mixed nets, env.step, terminal observations, auto re-deal) followed by the
move of the staged records into both players' ring (M_RL) and reservoir (M_SL) memories.
Workload (BASELINE.json configs[4] per GPU == configs[2] scaled to 1M games): 2^20 games per GPU,
eta 0.1, epsilon 0.06, Glorot-initialised 30-64-3 nets, synthetic Philox deals.  Weak scaling: games
shard by global id with no data-path collective.

Prints ONE JSON line (contract in the task statement).  `value` is device-timed with inputs resident in
HBM; `e2e` adds, per step, the pinned-host -> device copy of the four nets' weights, the sampling of a
256-row minibatch from all four memories and the device -> host read of those batches and the counters.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GAMES_PER_GPU = 1 << 20
T_PER_CALL = 8
ENV_T_PER_CALL = 32  # env-only K1 beside the headline: transitions per game and launch (trace planes: 402 MB)
ETA, EPS, SEED = 0.1, 0.06, 1234
RL_CAP, SL_CAP = 1 << 25, 1 << 23  # records per player (16 B each): one step's records never wrap the ring
BATCH = 256
BYTES_PER_TRANSITION = 33.6  # SURVEY 8d cfg 3: state 16 + RL record 16 + eta * SL record 16
ENV_BYTES_PER_TRANSITION = 28.0  # SURVEY 8d cfg 2: state 16 + 12-byte trace record


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML every 2 ms, nvidia-smi
    (one query per ~0.1 s) if NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.source = index, [], False, "nvidia-smi"
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def _nvml_row(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons")
                else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        bits = (0x8, 0x40, 0x20, 0x4)  # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap
        return [mhz, self.max_mhz] + ["Active" if r & b else "Not Active" for b in bits]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.rows.append(self._nvml_row())
                    time.sleep(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return None
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(str(r[2 + i]).lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "source": self.source}


def cpu_rollout_rate(threads, games_per_thread, t_steps, reps):
    """Oracle port of the same workload on `threads` host threads (independent game shards)."""
    import torch

    from oracle import orc

    # Keras-default Glorot nets, same formula as the GPU arm's (agent.py:101-103,110-112); the product
    # package is deliberately not imported on this arm
    g = torch.Generator().manual_seed(SEED)
    wt = torch.zeros((4, 2179), dtype=torch.float32)
    l1, l2 = (6.0 / 94) ** 0.5, (6.0 / 67) ** 0.5
    for k in range(4):
        wt[k, :1920] = (torch.rand(1920, generator=g) * 2 - 1) * l1
        wt[k, 1984:2176] = (torch.rand(192, generator=g) * 2 - 1) * l2
    w = wt.numpy()
    nets = orc.Nets([dict(W1=w[k, :1920].reshape(30, 64), b1=w[k, 1920:1984], W2=w[k, 1984:2176].reshape(64, 3),
                          b2=w[k, 2176:]) for k in range(4)])
    eta_u, eps_u = orc.u32_frac(ETA), orc.u32_frac(EPS)
    batches = []
    for t in range(threads):
        b = orc.NfspBatch(games_per_thread, SEED, game0=t * games_per_thread)
        b.reset(0, eta_u)
        batches.append(b)
    cap = 3 * games_per_thread * t_steps + 8

    def work(b, step0):
        b.rollout_act(step0, t_steps, nets, eta_u, eps_u, rec_cap=cap)

    def one(step0):
        ths = [threading.Thread(target=work, args=(b, step0)) for b in batches]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return time.perf_counter() - t0

    one(1)  # warm-up
    times = [one(1 + (i + 1) * t_steps) for i in range(reps)]
    trans = threads * games_per_thread * t_steps
    return trans, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc

    threads = max(1, os.cpu_count() or 1)
    gpt = 8192
    trans, times = cpu_rollout_rate(threads, gpt, T_PER_CALL, args.warmup + args.steps)
    times = times[args.warmup:]
    total = sum(times)
    value = trans * len(times) / total
    sample = "%d threads x %d games x %d decisions per step (oracle port of Agent.play + newenv + memories adds)" % (
        threads, gpt, T_PER_CALL)
    line = {"impl": "reference", "metric": "leduc_transitions_per_sec", "value": value, "unit": "transitions/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config():
    return {"workload": "nfsp_rollout: %d games/GPU x %d decisions per step, eta=%.2f eps=%.2f, 4 acting nets 30-64-3, "
                        "ring %d + reservoir %d records per player, sample %d" %
                        (GAMES_PER_GPU, T_PER_CALL, ETA, EPS, RL_CAP, SL_CAP, BATCH),
            "games_per_gpu": GAMES_PER_GPU, "decisions_per_step": T_PER_CALL, "l2": "flushed between timed steps",
            "baseline_config": "BASELINE.json configs[4] per GPU (configs[2] at 1M games)"}


def buffer_kernels(nfsp_b200, dev, ev, n_rec=1 << 23):
    """K3/K4/K5 beside the headline (SURVEY 8d cfg 4): insert batches of 2^23 staged records into a 2^24-slot
    ring / reservoir (HBM streaming / scatter), and the 256-row sample+gather (latency)."""
    import torch

    hbm, _ = peaks()
    recs = torch.randint(0, 1 << 30, (n_rec, 4), dtype=torch.int32, device=dev)
    out = {}
    ring = nfsp_b200.DeviceRing(1 << 24, 1, dev)
    res = nfsp_b200.DeviceReservoir(1 << 24, 2, dev)
    for name, mem, bytes_per in (("ring_insert", ring, 32.0), ("reservoir_insert", res, 40.0)):
        ms = []
        for k in range(8):
            cnt = torch.tensor([n_rec], dtype=torch.int32, device=dev)
            a, b = ev(), ev()
            a.record()
            mem.insert(recs, cnt)
            b.record()
            b.synchronize()
            if k >= 3:
                ms.append(a.elapsed_time(b))
        t = sum(ms) / len(ms)
        gbs = n_rec * bytes_per / (t * 1e-3) / 1e9
        out[name] = {"records": n_rec, "ms": t, "records_per_sec": n_rec / (t * 1e-3), "achieved_gbs": gbs,
                     "algorithmic_bytes_per_record": bytes_per, "frac_of_hbm_peak": gbs / hbm}
    # SURVEY 8d cfg 4 as stated: the 2 000 000-slot reservoir (stamps + payload stay in the 126 MB L2), batches of 2^20
    res2m = nfsp_b200.DeviceReservoir(2_000_000, 3, dev)
    n_small = 1 << 20
    ms = []
    for k in range(11):
        cnt = torch.tensor([n_small], dtype=torch.int32, device=dev)
        a, b = ev(), ev()
        a.record()
        res2m.insert(recs[:n_small], cnt)
        b.record()
        b.synchronize()
        if k >= 3:
            ms.append(a.elapsed_time(b))
    t = sum(ms) / len(ms)
    out["reservoir_insert_2M_slots"] = {"records": n_small, "ms": t, "records_per_sec": n_small / (t * 1e-3),
                                        "achieved_gbs": n_small * 40.0 / (t * 1e-3) / 1e9,
                                        "frac_of_hbm_peak": n_small * 40.0 / (t * 1e-3) / 1e9 / hbm,
                                        "note": "tickets 3M..11M into 2M slots: 18-66% of the records are accepted"}
    ms = []
    for k in range(13):
        a, b = ev(), ev()
        a.record()
        ring.sample(BATCH)
        res.sample(BATCH)
        b.record()
        b.synchronize()
        if k >= 3:
            ms.append(a.elapsed_time(b))
    out["sample_256_rl_plus_sl_us"] = 1e3 * sum(ms) / len(ms)
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import nfsp_b200
    from nfsp_b200 import sharding

    rank, world, local = sharding.env_rank_world()
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    game0, n = sharding.shard_games(GAMES_PER_GPU * world, rank, world)
    sp = nfsp_b200.SelfPlay(n, seed=SEED, game0=game0, device=dev, eta=ETA, epsilon=EPS, rl_capacity=RL_CAP,
                            sl_capacity=SL_CAP, max_steps_per_call=T_PER_CALL, variant=args.variant)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    w_host = sp.weights.cpu().pin_memory()
    stats_host = torch.empty(sp.stats.shape, dtype=sp.stats.dtype).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def timed_steps(k, e2e):
        gc.collect()
        gc.disable()  # a collector pause inside a host-driven step shows up as a straggler rank (max over ranks)
        try:
            return _timed_steps(k, e2e)
        finally:
            gc.enable()

    def _timed_steps(k, e2e):
        tot_ms, ker_ms = 0.0, 0.0
        for _ in range(k):
            flush_buf.zero_()  # L2 flush, outside the timed events
            a, b, c = ev(), ev(), ev()
            a.record()
            if e2e:
                sp.set_weights(w_host)  # pinned host -> device copy inside, then the kernels' weight images are rebuilt
            sp.rollout(T_PER_CALL, insert=False)
            b.record()
            sp.flush()
            if e2e:  # the learner's four minibatches and the counters come back to the host: one slab, one copy each
                sp.sample_minibatches(BATCH, to_host=True)
                stats_host.copy_(sp.stats, non_blocking=True)
            c.record()
            c.synchronize()
            tot_ms += a.elapsed_time(c)
            ker_ms += a.elapsed_time(b)
        return tot_ms, ker_ms

    timed_steps(args.warmup, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    tot_ms, ker_ms = timed_steps(args.steps, False)
    barrier()
    t_e2e, _ = timed_steps(max(1, min(args.warmup, 2)), True)  # e2e warm-up (pinned buffers, sample kernels)
    barrier()
    e2e_ms, _ = timed_steps(args.steps, True)
    barrier()
    sampler.stop_flag = True
    tot_ms = sharding.max_over_ranks(tot_ms, dev)
    e2e_ms = sharding.max_over_ranks(e2e_ms, dev)
    trans_all = GAMES_PER_GPU * world * T_PER_CALL * args.steps
    value = trans_all / (tot_ms * 1e-3)
    e2e_value = trans_all / (e2e_ms * 1e-3)
    st = sharding.allreduce_stats(sp.stats)

    # learner beside it (BASELINE configs[4]): update_strategy() of both agents = 8 SGD steps of the
    # four nets, each with ONE all-reduce of the flat gradient+stats buffer (NCCL when world > 1)
    from nfsp_b200.learner import Learner

    learner = Learner(sp, cfg=nfsp_b200.load_config(None))
    for _ in range(2):
        learner.update()
    barrier()
    a, b = ev(), ev()
    a.record()
    for _ in range(5):
        lstats = learner.update()
    b.record()
    b.synchronize()
    learner_ms = sharding.max_over_ranks(a.elapsed_time(b) / 5, dev)
    # the full self-play iteration of BASELINE configs[4]: rollout(8) + memory inserts + update_strategy() of both agents
    # (with its all-reduces when world > 1), nothing read back between iterations
    barrier()
    a, b = ev(), ev()
    a.record()
    for _ in range(5):
        sp.rollout(T_PER_CALL)
        learner.update(sync=False)
    b.record()
    b.synchronize()
    train_ms = sharding.max_over_ranks(a.elapsed_time(b) / 5, dev)
    # the same iteration with the update beside the next rollout (learner.PipelinedTrainer: acting nets one update behind)
    from nfsp_b200.learner import PipelinedTrainer

    pipe_ms = None
    if learner.one_launch:  # needs the one-launch fit (with several GPUs: peer memory); every rank takes the same branch
        trainer = PipelinedTrainer(sp, learner)
        for _ in range(3):
            trainer.step(T_PER_CALL)
        barrier()
        a, b = ev(), ev()
        a.record()
        for _ in range(10):
            trainer.step(T_PER_CALL)
        b.record()
        b.synchronize()
        pipe_ms = sharding.max_over_ranks(a.elapsed_time(b) / 10, dev)
        trainer.finish()
        torch.cuda.synchronize()

    # env-only K1 (BASELINE configs[1]) beside it, same games count, trace planes written
    env = nfsp_b200.BatchedNfspEnv(n, seed=SEED, game0=game0, device=dev)
    env.reset()
    tr = torch.empty((3, ENV_T_PER_CALL, n), dtype=torch.int32, device=dev)
    from nfsp_b200.batched import _ptr, _stream, check, lib

    def env_step():
        check(lib().nfsp_env_step(env._h, None, None, ENV_T_PER_CALL, 1, ETA, _ptr(tr), _stream(dev)))

    for _ in range(3):
        env_step()
    env_ms = 0.0
    for _ in range(args.steps):
        flush_buf.zero_()
        a, b = ev(), ev()
        a.record()
        env_step()
        b.record()
        b.synchronize()
        env_ms += a.elapsed_time(b)
    env_rate = n * ENV_T_PER_CALL * args.steps / (env_ms * 1e-3)

    # the README's env (leduc/env.py, ENV_LEGACY) beside it: README iterations of 2 transitions, records written
    leg = nfsp_b200.BatchedLegacyEnv(n, seed=SEED, game0=game0, device=dev)
    leg.reset()
    leg_iters = ENV_T_PER_CALL // 2  # as many transitions per game and launch as the NFSP env kernel above
    rec = torch.empty((3, leg_iters, n, 2), dtype=torch.int32, device=dev)

    def legacy_step():
        check(lib().nfsp_legacy_rollout(leg._h, None, leg_iters, _ptr(rec), _stream(dev)))

    for _ in range(3):
        legacy_step()
    leg_ms = 0.0
    for _ in range(args.steps):
        flush_buf.zero_()
        a, b = ev(), ev()
        a.record()
        legacy_step()
        b.record()
        b.synchronize()
        leg_ms += a.elapsed_time(b)
    legacy_rate = 2 * n * leg_iters * args.steps / (leg_ms * 1e-3)
    del rec, leg

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    buffers = buffer_kernels(nfsp_b200, dev, ev) if world == 1 else None
    # the other first-layer variant, kernel only, for the record
    other = "tcgen05" if args.variant in ("default", "cuda") else "cuda"
    other_ms = 0.0
    for k in range(3 + args.steps):
        flush_buf.zero_()
        a, b = ev(), ev()
        a.record()
        sp.rollout(T_PER_CALL, insert=False, variant=other)
        b.record()
        b.synchronize()
        sp.counts.zero_()
        if k >= 3:
            other_ms += a.elapsed_time(b)
    other_rate = n * T_PER_CALL * args.steps / (other_ms * 1e-3)
    # BASELINE configs[2] as stated: 64k parallel games (latency-bound: 2 048 blocks of 32 games for 4 736 warps)
    sp64 = nfsp_b200.SelfPlay(1 << 16, seed=SEED, device=dev, eta=ETA, epsilon=EPS, rl_capacity=200000,
                              sl_capacity=2000000, max_steps_per_call=T_PER_CALL)
    ms64 = 0.0
    for k in range(3 + args.steps):
        a, b = ev(), ev()
        a.record()
        sp64.rollout(T_PER_CALL)
        b.record()
        b.synchronize()
        if k >= 3:
            ms64 += a.elapsed_time(b)
    rate64 = (1 << 16) * T_PER_CALL * args.steps / (ms64 * 1e-3)
    del sp64
    hbm, which = peaks()
    kernel_rate = n * T_PER_CALL * args.steps / (ker_ms * 1e-3)  # this rank's rollout kernel alone
    achieved = kernel_rate * BYTES_PER_TRANSITION / 1e9
    roofline = {"bound": "hbm", "kernel": "rollout_tc_kernel" if args.variant == "tcgen05" else "rollout_kernel", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm, "traffic": None, "peak_source": which,
                "algorithmic_bytes_per_transition": BYTES_PER_TRANSITION,
                "kernel_ms_per_launch": ker_ms / args.steps, "useful_tflops": kernel_rate * 4224 / 1e12}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        tf = json.load(open(traffic_file))
        roofline["traffic"] = tf.get("rollout_kernel_bytes_per_launch")
        if args.variant != "tcgen05" and "rollout_kernel_ncu" in tf:
            # what really bounds the kernel (ncu, not measured in this run): each decision has to bring 2 x 256 B of
            # first-layer rows and 768 B of second-layer weights from shared memory into registers
            roofline["limiter"] = "shared-memory register-fill bandwidth, not HBM"
            roofline["ncu"] = tf["rollout_kernel_ncu"]
    cpu = None
    if world == 1:
        try:
            trans, times = cpu_rollout_rate(1, 65536, T_PER_CALL, 3)
            cpu = {"value": trans * len(times) / sum(times), "unit": "transitions/s", "cores": 1, "kind": "port",
                   "sample": "65536 games x %d decisions x %d passes, oracle port, 1 thread" % (T_PER_CALL, len(times))}
        except Exception as e:  # the oracle is a reported baseline, never the product
            cpu = {"value": None, "unit": "transitions/s", "cores": 1, "kind": "port", "sample": "failed: %r" % (e,)}
    h2d = int(w_host.numel() * 4)
    d2h = int(2 * BATCH * (30 + 3 + 1 + 30 + 1) * 4 + 2 * BATCH * 33 * 4 + sp.stats.numel() * 8)
    line = {"metric": "leduc_transitions_per_sec", "value": value, "unit": "transitions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "transitions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": 6 * args.steps,  # rollout + (ring insert, commit) + (reservoir stamp, write, commit), both players per launch
            "roofline": roofline, "cpu_baseline": cpu,
            "extra": {"env_only": {"kernel": "nfsp_step_fsm_kernel", "transitions_per_sec": env_rate,
                                   "achieved_gbs": env_rate * ENV_BYTES_PER_TRANSITION / 1e9,
                                   "frac_of_hbm_peak": env_rate * ENV_BYTES_PER_TRANSITION / 1e9 / hbm,
                                   "algorithmic_bytes_per_transition": ENV_BYTES_PER_TRANSITION,
                                   "transitions_per_launch": n * ENV_T_PER_CALL,
                                   # what the kernel really moves: the 16-byte state word once per launch, 12 bytes of trace per transition
                                   "moved_bytes_per_transition": 12.0 + 16.0 / ENV_T_PER_CALL,
                                   "moved_gbs": env_rate * (12.0 + 16.0 / ENV_T_PER_CALL) / 1e9,
                                   "moved_frac_of_hbm_peak": env_rate * (12.0 + 16.0 / ENV_T_PER_CALL) / 1e9 / hbm},
                      "env_only_legacy": {"kernel": "legacy_rollout_kernel", "transitions_per_sec": legacy_rate,
                                          "achieved_gbs": legacy_rate * 20.0 / 1e9, "frac_of_hbm_peak": legacy_rate * 20.0 / 1e9 / hbm,
                                          "algorithmic_bytes_per_transition": 20.0,
                                          "transitions_per_launch": 2 * n * leg_iters,
                                          "moved_bytes_per_transition": 12.0 + 8.0 / leg_iters,
                                          "moved_gbs": legacy_rate * (12.0 + 8.0 / leg_iters) / 1e9,
                                          "moved_frac_of_hbm_peak": legacy_rate * (12.0 + 8.0 / leg_iters) / 1e9 / hbm},
                      "variant": args.variant, "other_variant": {"name": other, "kernel_transitions_per_sec": other_rate,
                                                                 "kernel_ms_per_launch": other_ms / args.steps},
                      "rollout_64k_games": {"config": "BASELINE configs[2]: 65536 games, ring 200000 + reservoir 2000000, rollout(8) + memory inserts",
                                            "transitions_per_sec": rate64, "ms_per_step": ms64 / args.steps},
                      "buffers": buffers,
                      "training_step": {"what": "rollout(8) + memory inserts + Learner.update(sync=False) per iteration (BASELINE configs[4])",
                                        "ms": train_ms, "transitions_per_sec": GAMES_PER_GPU * world * T_PER_CALL / (train_ms * 1e-3),
                                        "pipelined_ms": pipe_ms,
                                        "pipelined_transitions_per_sec": (GAMES_PER_GPU * world * T_PER_CALL / (pipe_ms * 1e-3)) if pipe_ms else None,
                                        "pipelined": "update j beside rollout j+1 (acting nets one update behind), 4 SMs left to the learner"},
                      "learner": {"update_ms": learner_ms, "sgd_steps_per_update": 8, "allreduce_floats": 4 * 2179 + 8,
                                  "exploitability_proxy": lstats.get("exploitability"), "trained_mask": lstats.get("trained")},
                      "hands": int(st[10]), "transitions_counted": int(st[11]), "records_dropped": int(st[12])}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="default", choices=["default", "cuda", "tcgen05"],
                    help="first layer of the acting nets: CUDA-core row sums or tcgen05 tensor-core tiles")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
