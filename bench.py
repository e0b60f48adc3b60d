#!/usr/bin/env python
"""Benchmark of the ABM hot path (visual field + flocking update), BASELINE.json metric:
agent-steps/sec at N agents x R replicates on 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE configs[3] -- visual flocking, 1024 agents x 1024
replicates PER GPU (weak scaling: replicates are independent, no data-path collective),
R = 1200, full FOV, walls, reference disc initial condition (vf_sims.py:212-228),
parameters of VFExp4c (GAM .1, V0 1, ALP0 1, ALP1 .09, BET0 1, BET1 .09), arena
ceil(900 * sqrt(N/100)) = 2880 px, radius 10.  A "step" = one fused kernel launch that
advances all agents of all replicates of the rank by one time step.

One JSON line is printed by rank 0 (see the driver contract in the task statement).
`--impl reference` times the CPU port of the reference's own algorithm (oracle/literal.py;
the reference is pure Python and /root/reference does not exist on the GPU box).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_AGENTS = int(os.environ.get("ABM_BENCH_AGENTS", 1024))
N_REPLICATES = int(os.environ.get("ABM_BENCH_REPLICATES", 1024))
R = 1200
RADIUS = 10.0
PAD = 30.0
PARAMS = dict(GAM=0.1, V0=1.0, ALP0=1.0, ALP1=0.09, BET0=1.0, BET1=0.09)
# dram__bytes_read.sum + dram__bytes_write.sum of one step-kernel launch at the default workload, from the committed
# `ncu --set full` captures (profiles/r1_*_ncu_summary.md); None for any other workload size or kernel
NCU_DRAM_BYTES_PER_LAUNCH = {"abm::vf_step_kernel": 55.2e6, "abm::vf_step_sym_kernel": 26.2e6} \
    if (N_AGENTS, N_REPLICATES) == (1024, 1024) else {}
METRIC = "agent-steps/sec (visual field + flocking update)"
UNIT = "agent-steps/s"


def arena_side(n_agents):
    return float(math.ceil(900.0 * math.sqrt(n_agents / 100.0)))


def synthetic_state(n_rep, n_agents, first_replicate=0):
    """Reference disc initial condition (vf_sims.py:216-221), seeds default_rng(1234 + replicate)
    (SURVEY 8d).  Returns float32 (n_rep, n_agents) arrays x, y, theta, vel."""
    W = arena_side(n_agents)
    x = np.empty((n_rep, n_agents), np.float32)
    y = np.empty_like(x)
    th = np.empty_like(x)
    for b in range(n_rep):
        rng = np.random.default_rng(1234 + first_replicate + b)
        orient = rng.uniform(0, 2 * np.pi, n_agents)
        dist = rng.uniform(0, 1, n_agents) * (W / 2 - 2 * PAD - RADIUS)
        x[b] = np.cos(orient) * dist + W / 2
        y[b] = np.sin(orient) * dist + W / 2
        th[b] = orient
    return x, y, th, np.zeros_like(x)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU phases run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 8:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[1]) for s in self.samples)
        reasons = []
        for idx, name in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"),
                          (7, "sw_power_cap")):
            if any(s[idx].lower().startswith("active") for s in self.samples):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][2]), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(s[3]) for s in self.samples)}


def visible_pair_fraction(x, y, r=RADIUS):
    """Fraction of ordered pairs whose half width is >= 1 bin: d <= r / tan(2pi/R)."""
    import torch
    lim2 = (r / math.tan(2 * math.pi / R)) ** 2
    cx, cy = x + r, y + r
    d2 = (cx[:, :, None] - cx[:, None, :]) ** 2 + (cy[:, :, None] - cy[:, None, :]) ** 2
    n = x.shape[1]
    vis = (d2 <= lim2).sum().item() - x.shape[0] * n
    return vis / float(x.shape[0] * n * (n - 1))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


_CPU = {}     # state of the CPU arm, built once per process and inherited by forked workers


def _cpu_setup(n_agents, seed_replicate=0):
    """The frozen snapshot the CPU arm works on: replicate `seed_replicate` of the workload's initial state.  With the
    reference's own modules at hand (oracle/_ref, copied by oracle/build_ref.py; /root/reference in the build container)
    the arm is the UNMODIFIED `VFAgent.update` (vf_agent.py:52-80) on real VFAgent objects (pygame stubbed: drawing is
    free, which favours the reference); otherwise the float64 port oracle/literal.py."""
    from oracle import ref_shim, restate as rs
    x, y, th, v = synthetic_state(1, n_agents, seed_replicate)
    x64, y64, th64, v64 = (a[0].astype(np.float64) for a in (x, y, th, v))
    W = arena_side(n_agents)
    _CPU.clear()
    _CPU.update(kind="port", x=x64, y=y64, th=th64, v=v64, rad=np.full(n_agents, RADIUS),
                cfg=rs.VFConfig(R=R, width=W, height=W, **PARAMS))
    if ref_shim.reference_available() and os.environ.get("ABM_BENCH_CPU_KIND", "reference") != "port":
        _CPU["agents"] = ref_shim.make_vf_agents(x64, y64, th64, v64, int(RADIUS), R=R, width=W, height=W,
                                                 boundary="walls", params=PARAMS)
        _CPU["kind"] = "reference"
    return _CPU["kind"]


def _cpu_chunk(idx):
    """Update the focal agents `idx` from the frozen snapshot (every one sees the un-updated others)."""
    if _CPU["kind"] == "reference":
        import copy
        agents = _CPU["agents"]
        for i in idx:
            a = copy.copy(agents[i])              # update() writes the focal agent alone: a shallow copy with its own
            a.position = np.array(a.position)     # position array keeps the snapshot frozen (a deep copy would also
            a.update(agents)                      # duplicate the agent's arena-sized line map, 68 MB at this size)
    else:
        from oracle import literal
        c = _CPU
        for i in idx:
            literal.agent_update(i, c["x"], c["y"], c["th"], c["v"], c["rad"], c["cfg"])
    return len(idx)


def cpu_rate(n_focal, procs=1, pool=None):
    """agent-steps/s of the CPU arm set up by _cpu_setup: `n_focal` focal agents against the full neighbour set."""
    idx = list(range(n_focal))
    if procs <= 1:
        t0 = time.perf_counter()
        _cpu_chunk(idx)
        dt = time.perf_counter() - t0
    else:
        chunks = [idx[p::procs] for p in range(procs)]
        t0 = time.perf_counter()
        pool.map(_cpu_chunk, chunks)
        dt = time.perf_counter() - t0
    return n_focal / dt, dt


def cpu_sample_text(kind, n_focal, n_agents, procs, dt=None):
    what = ("the UNMODIFIED reference VFAgent.update (vf_agent.py:52-80, modules copied to oracle/_ref by "
            "oracle/build_ref.py; pygame stubbed, so drawing is free)") if kind == "reference" else \
           "oracle/literal.py, float64 port of VFAgent.update (same per-pair loop and arg-min scans as the reference)"
    t = f" ({dt:.1f} s)" if dt is not None else ""
    return (f"{n_focal} focal agents x {n_agents - 1} neighbours of replicate 0 per step{t}, frozen snapshot, "
            f"{procs} process(es) (one per host core); {what}")


def run_reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path on all host cores (one process per core: the reference's own
    way of using a node, HPC_batch_run.sh / README.md:302-324)."""
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    n_focal = max(procs * 2, int(os.environ.get("ABM_BENCH_CPU_FOCAL", 16 * procs)))
    kind = _cpu_setup(N_AGENTS)
    import multiprocessing as mp
    rates = []
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_cpu_chunk, [[]] * procs)                          # start the workers
        for s in range(args.warmup + args.steps):
            rate, dt = cpu_rate(n_focal, procs=procs, pool=pool)
            if s >= args.warmup:
                rates.append((rate, dt))
    value = float(np.mean([r for r, _ in rates]))
    ms = float(np.mean([d for _, d in rates]) * 1e3)
    sample = cpu_sample_text(kind, n_focal, N_AGENTS, procs)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world):
    tag = "BASELINE configs[3]: " if (N_AGENTS, N_REPLICATES) == (1024, 1024) else "(size overridden by ABM_BENCH_*) "
    return {"workload": f"{tag}visual flocking, {N_AGENTS} agents x {N_REPLICATES} replicates per GPU, R={R}",
            "n_agents": N_AGENTS, "replicates_per_gpu": N_REPLICATES, "replicates_total": N_REPLICATES * world,
            "fov_resolution": R, "boundary": "walls", "arena_px": arena_side(N_AGENTS), "agent_radius": RADIUS,
            "params": PARAMS, "parallelism": f"replicate-sharded x{world}, no data-path collective",
            "l2": "L2 flushed (256 MiB memset) between timed steps; flush excluded from the timing",
            "update": "synchronous (Jacobi) step from a frozen snapshot; every unordered pair evaluated once (fp32 + "
                      "32-bit binary angles), near-tie pairs re-evaluated in fp64, epilogue in fp64"}


def parity_check(VFEngine, local_rank, x, y, th, v, rad, n_check=3):
    """Correctness evidence in the bench record itself: ONE step of the benchmark's own kernel (the engine picks it from
    the workload's size exactly as in the timed arm) from the workload's initial state, with fields and terms kept,
    compared with the float64 oracle (oracle/restate.vf_step_frozen: the reference's VFAgent.update per agent from the
    frozen snapshot, vf_agent.py:52-80 / vf_supcalc.py:20-138) for EVERY agent of `n_check` replicates.  Outside all
    timed regions.  The checker, not the thing measured."""
    from oracle import restate as rs
    B, N = x.shape
    W = arena_side(N)
    eng = VFEngine(B, N, resolution=R, width=W, height=W, device=local_rank, keep_fields=True, keep_terms=True)
    eng.set_params(**PARAMS)
    eng.set_state(x, y, th, v, rad)
    eng.step(1)
    kernel = eng.last_kernel()
    fields, terms, st, ctr = eng.fields_packed(), eng.terms(), eng.get_state(), eng.counters()
    eng.close()
    cfg = rs.VFConfig(R=R, width=W, height=W, **PARAMS)
    reps = sorted({0, B // 2, B - 1})[:n_check]
    bits = 0
    rel_terms = rel_state = 0.0
    edges = 0
    for b in reps:
        ref = rs.vf_step_frozen(x[b], y[b], th[b], v[b], RADIUS, cfg)
        want = rs.pack_bits(ref["rows"][:, ::-1])                       # stored (flipped) order
        diff = np.bitwise_xor(want, fields[b])
        bits += int(np.unpackbits(diff.view(np.uint8)).sum())
        edges += int((ref["rows"] != np.roll(ref["rows"], 1, axis=1)).sum())
        rel_terms = max(rel_terms, float(np.max(np.abs(terms[b] - ref["terms"]) / np.maximum(np.abs(ref["terms"]), 1e-6))))
        for k in ("x", "y", "theta", "vel"):
            rel_state = max(rel_state, float(np.max(np.abs(st[k][b] - ref[k]) / np.maximum(np.abs(ref[k]), 1e-3))))
    return {"kernel": kernel, "replicates": reps, "agents": len(reps) * N, "bins": len(reps) * N * R,
            "field_bits_differ": bits, "max_rel_terms": rel_terms, "max_rel_state": rel_state,
            "fp64_pairs": ctr["fp64_pairs"], "fp32_fp64_index_differ": ctr["fp32_fp64_differ"],
            "edges_per_agent": edges / float(len(reps) * N),
            "oracle": "oracle/restate.vf_step_frozen (float64 restatement pinned to the reference's golden vectors and "
                      "to fixtures of the unmodified reference), every agent of the listed replicates, one step from "
                      "the workload's initial state",
            "tolerance": "fields bit-exact; terms / state 1e-5 relative (north_star)",
            "ok": bool(bits == 0 and rel_terms < 1e-5 and rel_state < 1e-5)}


def swarm_arm(VFEngine, local_rank, rank, world, dist, steps=20):
    """BASELINE configs[4]: ONE swarm of 65 536 agents (visual flocking, R = 1200, arena 23 040 px, walls, disc initial
    condition), agent tiles sharded across the ranks; the only data exchanged per step are the 16-byte records
    (1 MiB), stored by the step kernel straight into the peers' tables over NVLink (abm_b200/multigpu.TiledSwarm,
    fused=True).  STRONG scaling: total work fixed.  Reports ms per step (CUDA events, max over ranks),
    the same swarm on ONE GPU timed by rank 0 in the same process, whether the sharded run ends in the single-GPU run's
    exact state, and an oracle check of sampled agents of the 65 536-agent scene."""
    import torch
    from abm_b200.multigpu import TiledSwarm
    from oracle import restate as rs
    N = int(os.environ.get("ABM_BENCH_SWARM_AGENTS", 65536))
    W = arena_side(N)
    x, y, th, v = synthetic_state(1, N)
    kw = dict(resolution=R, width=W, height=W, boundary="walls")
    warm = 4
    out = {"n_agents": N, "arena_px": W, "steps": steps, "boundary": "walls"}

    def timed(step_fn):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record(); step_fn(steps); e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    ref_state = None
    if rank == 0:      # the same swarm on one GPU: oracle spot check after one step, reference state, 1-GPU time
        e1 = VFEngine(1, N, device=local_rank, keep_fields=True, keep_terms=True, **kw)
        e1.set_params(**PARAMS); e1.set_state(x, y, th, v, RADIUS); e1.step(1)
        sample = sorted(np.random.default_rng(5).choice(N, 8, replace=False).tolist())
        cfg = rs.VFConfig(R=R, width=W, height=W, **PARAMS)
        ref = rs.vf_step_frozen(x[0], y[0], th[0], v[0], RADIUS, cfg, agents=sample)
        f1, t1, s1 = e1.fields()[0], e1.terms()[0], e1.get_state()
        bits = int((f1[sample] != ref["rows"][sample][:, ::-1]).sum())
        rel = float(np.max(np.abs(t1[sample] - ref["terms"][sample]) / np.maximum(np.abs(ref["terms"][sample]), 1e-6)))
        for k in ("x", "y", "theta", "vel"):
            rel = max(rel, float(np.max(np.abs(s1[k][0][sample] - ref[k][sample]) / np.maximum(np.abs(ref[k][sample]), 1e-3))))
        out["oracle_check"] = {"agents": sample, "field_bits_differ": bits, "max_rel": rel,
                               "ok": bool(bits == 0 and rel < 1e-5), "kernel": e1.last_kernel()}
        e1.step(warm - 1)
        ref_state = e1.get_state()
    if world == 1:
        ms1 = timed(e1.step)
        out.update(ms_per_step=ms1, ms_per_step_1gpu=ms1, exchange="none (one GPU)", strong_scaling_eff=1.0,
                   bit_identical_to_1gpu=True, agent_steps_per_s=N / ms1 * 1e3)
        e1.close()
        return out
    # rank 0 times the single-GPU swarm while the other ranks wait at the barrier inside timed()
    t1gpu = torch.zeros(1, dtype=torch.float64, device="cuda")
    if rank == 0:
        e0 = torch.cuda.Event(enable_timing=True); e1e = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); e1.step(steps); e1e.record(); torch.cuda.synchronize()
        t1gpu[0] = e0.elapsed_time(e1e) / steps
        e1.close()
    dist.broadcast(t1gpu, 0)
    swarm = TiledSwarm(N, fused=True, **kw)
    swarm.set_params(**PARAMS)
    swarm.set_state(x, y, th, v, RADIUS)
    swarm.step(warm)
    got = swarm.get_state()
    same = True
    if rank == 0:
        same = all(np.array_equal(got[k], ref_state[k][0]) for k in ("x", "y", "theta", "vel"))
    ms = torch.tensor([timed(swarm.step)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    flag = torch.tensor([1 if same else 0], device="cuda"); dist.broadcast(flag, 0)
    swarm.engine.close()
    ms, t1 = float(ms.item()), float(t1gpu.item())
    out.update(ms_per_step=ms, ms_per_step_1gpu=t1, exchange="peer-store" if swarm.fused else "nccl",
               tiles="cyclic blocks of 128 slots" if swarm.cyclic else "contiguous", strong_scaling_eff=t1 / (world * ms),
               bit_identical_to_1gpu=bool(flag.item()), agent_steps_per_s=N / ms * 1e3)
    return out


def other_configs_arm(VFEngine, local_rank, peaks):
    """The other BASELINE configs on one GPU (SURVEY 8d lists all five; the headline line is configs[3], the swarm arm
    configs[4]): configs[0] foraging N = 10, 3 patches, T = 1000; configs[1] visual flocking N = 100, one run;
    configs[2] foraging sweep 1024 replicates x 50 agents with occlusion and collisions -- device-resident steps timed
    with CUDA events, plus for configs[2] an oracle check of the agent phase and its issue roofline."""
    import torch
    from abm_b200 import BaseEngine
    from oracle import restate_base as rb
    out = {}

    def timed(fn, n):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    prm = dict(Eps_u=1.0, F_N=0.5, F_R=0.5, exp_vel_max=3.0, exp_theta_min=-0.5, exp_theta_max=0.5,
               reloc_theta_max=1.8, exp_stop_ratio=0.175)
    # ---- configs[2]: figExp3BN50PatchyCollOcc-like sweep (DEC_EPSW one value per replicate) ----
    B, N, P, W = 1024, 50, 3, 500.0
    rng = np.random.default_rng(3)
    eng = BaseEngine(B, N, P, resolution=R, width=W, height=W, visual_exclusion=True, collide_agents=True,
                     ghost_mode=False, seed=9, keep_fields=True, device=local_rank)
    eng.set_params(Eps_w=np.tile(np.array([0, .25, .5, .75, 1, 2, 5, 3], np.float64), B // 8), **prm)
    eng.set_agents(x=rng.integers(20, 520, (B, N)), y=rng.integers(20, 520, (B, N)), theta=rng.uniform(0, 2 * np.pi, (B, N)))
    eng.set_patches(x=rng.integers(60, 400, (B, P)), y=rng.integers(60, 400, (B, P)), radius=np.full((B, P), 30.0),
                    left=np.full((B, P), 200.0), quality=np.full((B, P), 0.25), id=np.tile(np.arange(P), (B, 1)))
    eng.step(100)
    l0 = eng.counters()["launches"]
    ms = timed(eng.step, 200)
    launches = eng.counters()["launches"] - l0
    st = eng.get_agents()
    # parity of the agent phase from the state the timed run ended in: two replicates against the oracle
    dth = np.asarray(rng.uniform(-0.5, 0.5, (B, N)), np.float32)
    eng.step(1, inject_dtheta=dth, phases=2)
    got, fields = eng.get_agents(), eng.fields()
    bits, rel = 0, 0.0
    for b in (0, B - 1):
        cfg = rb.BaseConfig(R=R, width=W, height=W, visual_exclusion=True, Eps_w=float([0, .25, .5, .75, 1, 2, 5, 3][b % 8]), **prm)
        nov = ((st["novelty"][b][:, None] >> np.arange(cfg.Tau, dtype=np.uint32)) & 1).astype(float)
        s0 = dict(x=st["x"][b].astype(float), y=st["y"][b].astype(float), theta=st["theta"][b].astype(float),
                  vel=st["vel"][b].astype(float), radius=10.0, w=st["w"][b].astype(float), u=st["u"][b].astype(float),
                  novelty=nov, env_status=st["env_status"][b], override=st["override_mode"][b], mode=st["mode"][b],
                  patch_id=st["patch_id"][b], collected=st["collected"][b].astype(float),
                  collected_before=st["collected_before"][b].astype(float))
        ref = rb.base_step_frozen(s0, cfg, dth[b].astype(np.float64))
        bits += int((fields[b] != ref["fields"]).sum())
        for k, g in dict(x="x", y="y", theta="theta", vel="vel", w="w", u="u").items():
            rel = max(rel, float(np.max(np.abs(got[g][b] - ref[k]) / np.maximum(np.abs(ref[k]), 1e-3))))
    n_expl = float((st["override_mode"] == 1).sum()) / B
    ops = 48.0 * B * N * (N - 1) + 334.0 * B * N                      # the C3 roofline formula
    ops_occl = 8.0 * B * N * n_expl * (N - 1) / 2.0                   # OPS_OCCL estimate: every cue against half the others
    props = torch.cuda.get_device_properties(local_rank)
    peak = props.multi_processor_count * 128 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    out["c3"] = {"workload": "BASELINE configs[2]: foraging sweep, 1024 replicates x 50 agents, 3 patches, R=1200, occlusion + "
                             "collisions, one DEC_EPSW per replicate", "ms_per_step": ms, "agent_steps_per_s": B * N / ms * 1e3,
                 "gpu_launches_per_step": launches / 200.0, "kernel": "abm::base_step_kernel (collisions, environment and "
                 "agent phases of a replicate in one CTA)",
                 "roofline": {"bound": "fp32 issue", "achieved": (ops + ops_occl) / (ms * 1e-3) / 1e12, "peak": peak / 1e12,
                              "unit": "Tlaneop/s", "frac": (ops + ops_occl) / (ms * 1e-3) / peak, "ops_pairs_agents": ops,
                              "ops_occlusion_estimate": ops_occl, "exploiters_per_replicate": n_expl},
                 "parity": {"replicates": [0, B - 1], "agents": 2 * N, "field_bits_differ": bits, "max_rel_state": rel,
                            "ok": bool(bits == 0 and rel < 1e-5), "oracle": "oracle/restate_base.base_step_frozen"}}
    eng.close()
    # ---- configs[0]: one foraging run, N = 10, 3 patches, T = 1000 (every step of the run inside ONE launch) ----
    eng = BaseEngine(1, 10, 3, resolution=R, width=W, height=W, visual_exclusion=True, seed=4, device=local_rank)
    eng.set_params(Eps_w=2.0, **prm)
    eng.set_agents(x=rng.integers(20, 520, (1, 10)), y=rng.integers(20, 520, (1, 10)), theta=rng.uniform(0, 2 * np.pi, (1, 10)))
    eng.set_patches(x=rng.integers(60, 400, (1, 3)), y=rng.integers(60, 400, (1, 3)), radius=np.full((1, 3), 30.0),
                    left=np.full((1, 3), 200.0), quality=np.full((1, 3), 0.25), id=np.arange(3)[None])
    eng.step(50)
    l0 = eng.counters()["launches"]
    ms = timed(eng.step, 1000)
    out["c1"] = {"workload": "BASELINE configs[0]: foraging, N=10, 3 patches, R=1200, T=1000", "us_per_step": ms * 1e3,
                 "agent_steps_per_s": 10 / ms * 1e3, "gpu_launches_for_1000_steps": eng.counters()["launches"] - l0}
    eng.close()
    # ---- configs[1]: one visual-flocking run of 100 agents ----
    N2 = 100
    x, y, th, v = synthetic_state(1, N2)
    W2 = arena_side(N2)
    e2 = VFEngine(1, N2, resolution=R, width=W2, height=W2, device=local_rank)
    e2.set_params(**PARAMS); e2.set_state(x, y, th, v, RADIUS); e2.step(20)
    ms = timed(e2.step, 2000)
    out["c2"] = {"workload": "BASELINE configs[1]: visual flocking, N=100, one run, R=1200", "us_per_step": ms * 1e3,
                 "agent_steps_per_s": N2 / ms * 1e3, "kernel": e2.last_kernel()}
    e2.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference_arm(args, rank, world)
        return
    args.steps = 100 if args.steps is None else args.steps
    args.warmup = 3 if args.warmup is None else max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (abm_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner (NCCL_DEBUG=VERSION or WARN; the GPU boxes set VERSION) with printf on STDOUT,
        # where the one JSON line goes, when the communicator is created: file descriptor 1 points to stderr while
        # that happens (NCCL_DEBUG_FILE does not catch the banner)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from abm_b200 import VFEngine

    B, N = N_REPLICATES, N_AGENTS
    W = arena_side(N)
    x, y, th, v = synthetic_state(B, N, first_replicate=rank * B)
    rad = np.full((B, N), RADIUS, np.float32)
    eng = VFEngine(B, N, resolution=R, width=W, height=W, device=local_rank)
    eng.set_params(**PARAMS)
    dev = [torch.from_numpy(a).cuda() for a in (x, y, th, v, rad)]
    eng.set_state(*dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    with sampler:
        # ---- kernel-only arm: state resident in HBM ----
        for _ in range(args.warmup):
            eng.step(1)
        sx = torch.empty(B, N, device="cuda"); sy = torch.empty(B, N, device="cuda")
        eng.get_state({"x": sx, "y": sy})
        torch.cuda.synchronize()
        vis0 = visible_pair_fraction(sx[:16], sy[:16])
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        barrier()
        t_wall0 = time.perf_counter()
        for s in range(args.steps):
            flush.zero_()                      # L2 flush, outside the timed interval
            starts[s].record()
            eng.step(1)
            ends[s].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        step_ms = [a.elapsed_time(b) for a, b in zip(starts, ends)]
        total_ms = float(sum(step_ms))
        eng.get_state({"x": sx, "y": sy})
        torch.cuda.synchronize()
        vis1 = visible_pair_fraction(sx[:16], sy[:16])
        launches = args.steps

        # ---- end-to-end arm: host buffers in, host buffers out, every step ----
        # The way a sweep uses the plugin: THREE replicate batches of the workload's size in flight, each with its own
        # engine, stream and pair of pinned host buffers (the state a step returns is that batch's next input).  Every
        # batch-step copies its inputs host -> device, steps, and copies its result device -> host inside the timed
        # region; the copy engines move two batches while the SMs step the third (ABM_HOST_PINNED_ASYNC calls).
        hr = torch.from_numpy(rad).pin_memory().numpy()
        NB = int(os.environ.get('ABM_E2E_BATCHES', '3'))
        e2e_api = os.environ.get('ABM_E2E_API', 'step_host')
        # ONE interleaved pinned buffer (x, y, theta, vel per agent) per direction and batch: one copy each way per step
        packed0 = np.ascontiguousarray(np.stack([x, y, th, v], axis=-1))

        def e2e_arm(nb):
            extra = [VFEngine(B, N, resolution=R, width=W, height=W, device=local_rank) for _ in range(nb - 1)]
            for e_ in extra:
                e_.set_params(**PARAMS)
            engs = [eng] + extra
            streams = [torch.cuda.Stream() for _ in range(nb)]
            hosts = [[torch.from_numpy(packed0.copy()).pin_memory().numpy() for _ in range(2)] for _ in range(nb)]
            n_steps = nb * max(2, min(args.steps, 60) // nb)            # batch-steps, round robin over the batches
            cur = [0] * nb

            def batch_step(k, first=False):
                with torch.cuda.stream(streams[k]):
                    # H2D of the step's inputs; the radii are constants of the run, uploaded with the first call
                    if first or e2e_api != 'step_host':
                        engs[k].set_state_packed(hosts[k][cur[k]], hr if first else None, nonblocking=True)
                        engs[k].step(1)
                        engs[k].get_state_packed(hosts[k][cur[k] ^ 1], nonblocking=True)   # D2H of the step's result
                    else:   # the same three in one call; replicate chunks pipelined over the engine's copy streams
                        engs[k].step_host(hosts[k][cur[k]], hosts[k][cur[k] ^ 1], 1)
                    cur[k] ^= 1

            torch.cuda.synchronize()
            for k in range(nb):
                batch_step(k, first=True); batch_step(k)
            torch.cuda.synchronize()
            barrier()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(nb):
                streams[k].wait_event(e0)
            t_host0 = time.perf_counter()
            for it in range(n_steps):
                batch_step(it % nb)
            t_enq = time.perf_counter() - t_host0
            for k in range(nb):
                done = torch.cuda.Event(); done.record(streams[k]); torch.cuda.current_stream().wait_event(done)
            e1.record()
            barrier()
            for j in range(nb):                                            # every batch came back whole
                assert np.isfinite(hosts[j][cur[j]]).all()
            for e_ in extra:
                e_.close()
            return e0.elapsed_time(e1), n_steps, t_enq

        e2e1_ms, e2e1_steps, _ = e2e_arm(1)                              # one batch in flight (reported beside the headline)
        e2e_ms, e2e_steps, t_host_enqueue = e2e_arm(NB) if NB > 1 else (e2e1_ms, e2e1_steps, _)
    clocks = sampler.summary()
    counters = eng.counters()
    timed_kernel = eng.last_kernel()
    kernel_stats = eng.kernel_stats()

    # ---- parity of the benchmark's own kernel against the oracle (every rank on its own replicates) ----
    par = parity_check(VFEngine, local_rank, x, y, th, v, rad)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, par)
        par = dict(gathered[0], agents=sum(g["agents"] for g in gathered), bins=sum(g["bins"] for g in gathered),
                   field_bits_differ=sum(g["field_bits_differ"] for g in gathered),
                   max_rel_terms=max(g["max_rel_terms"] for g in gathered),
                   max_rel_state=max(g["max_rel_state"] for g in gathered),
                   fp64_pairs=sum(g["fp64_pairs"] for g in gathered),
                   fp32_fp64_index_differ=sum(g["fp32_fp64_index_differ"] for g in gathered),
                   ok=all(g["ok"] for g in gathered), ranks=world,
                   kernels=sorted({g["kernel"] for g in gathered}))
    par["same_kernel_as_timed"] = par["kernel"] == timed_kernel
    eng.close()

    # the headline's times, max over ranks (before the auxiliary arms: they cannot cost the line its numbers)
    t = torch.tensor([total_ms, e2e_ms, e2e1_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, e2e1_ms = (float(v) for v in t.tolist())

    # ---- BASELINE configs[4]: one swarm of 65 536 agents, agent tiles sharded across the ranks ----
    # (the auxiliary arms run after the headline's timed regions; an exception in one of them is reported in its object and
    # on stderr instead of costing the run its JSON line)
    swarm = None
    if os.environ.get("ABM_BENCH_SWARM", "1") != "0":
        try:
            swarm = swarm_arm(VFEngine, local_rank, rank, world, dist)
        except Exception as exc:                                      # noqa: BLE001
            print(f"bench.py: swarm arm failed on rank {rank}: {exc!r}", file=sys.stderr, flush=True)
            if world > 1:       # the other ranks may be inside the arm's exchange: fail fast rather than leave them waiting
                raise
            swarm = {"error": repr(exc)}
    others = None
    if world == 1 and os.environ.get("ABM_BENCH_OTHER_CONFIGS", "1") != "0":
        try:
            others = other_configs_arm(VFEngine, local_rank, measured_peaks()[0])
        except Exception as exc:                                      # noqa: BLE001
            print(f"bench.py: other_configs arm failed: {exc!r}", file=sys.stderr, flush=True)
            others = {"error": repr(exc)}

    agents_total = B * N * world
    value = agents_total * args.steps / (total_ms * 1e-3)
    e2e_value = agents_total * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        props = torch.cuda.get_device_properties(local_rank)
        sms = props.multi_processor_count
        vis = 0.5 * (vis0 + vis1)
        pairs = B * N * (N - 1)
        # OPS_PAIR = 6 * P_all + 42 * P_vis ; OPS_AGENT = 8 * ceil(R/32) + 12 * n_edges + 30
        n_edges = float(par.get("edges_per_agent", 0.0))              # measured on the parity step (workload's first step)
        ops_launch = 6.0 * pairs + 42.0 * pairs * vis + B * N * (8 * ((R + 31) // 32) + 12.0 * n_edges + 30)
        avg_s = total_ms * 1e-3 / args.steps
        sm_clock = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
        peak_ops = sms * 128 * sm_clock
        bytes_launch = 44.0 * B * N
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(4 * 4 * B * N),
                    "d2h_bytes_per_step": int(4 * 4 * B * N), "steps": e2e_steps,
                    "host_enqueue_s": t_host_enqueue, "device_s": e2e_ms * 1e-3,
                    "api": ("VFEngine.step_host (abm_vf_step_host: upload, step, download in one non-blocking call; replicate "
                            "chunks of whole CTA waves, copies on the engine's two copy streams overlap the other chunks' steps)"
                            if e2e_api == 'step_host' else
                            "VFEngine.set_state_packed -> step -> get_state_packed (abm_set_state_packed / abm_vf_step / "
                            "abm_get_state_packed, ABM_HOST_PINNED_ASYNC)") + ": one pinned (x, y, theta, vel) array per direction",
                    "pipeline": "%d batches in flight (one engine, stream and pair of pinned buffers each): copies overlap the "
                                "other batches' steps, and their CTAs fill each other's tail waves (which is why this can exceed "
                                "`value`, measured on ONE batch with an L2 flush between steps)" % NB,
                    "one_batch_in_flight": agents_total * e2e1_steps / (e2e1_ms * 1e-3)},
            "gpu_launches": launches,
            "roofline": {"bound": "fp32", "achieved": ops_launch / avg_s / 1e12, "peak": peak_ops / 1e12,
                         "unit": "Tlaneop/s", "frac": ops_launch / avg_s / peak_ops,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(timed_kernel), "kernel": timed_kernel,
                         "kernel_launches_by_variant": kernel_stats,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, committed `ncu --set full` "
                                           "capture under profiles/r2_vf_step_sym_ncu_summary.md (not measured in this run)",
                         "algorithmic_ops_per_launch": ops_launch,
                         "visible_pair_fraction": vis, "peak_source": f"{sms} SMs x 128 lanes x "
                         f"{sm_clock / 1e6:.0f} MHz ({peak_kind} sm_max_mhz)",
                         "hbm": {"achieved": bytes_launch / avg_s / 1e9, "peak": peaks.get("hbm_gbs"),
                                 "unit": "GB/s", "frac": bytes_launch / avg_s / 1e9 / peaks.get("hbm_gbs", 6650.0),
                                 "algorithmic_bytes_per_launch": bytes_launch, "peak_source": peak_kind}},
            "pairs_per_sec": pairs * world * args.steps / (total_ms * 1e-3),
            "fp64_pairs_fraction": counters["fp64_pairs"] / float(pairs * counters["launches"]),
            "wall_s_timed_region": t_wall,
            "parity": par,
            "swarm": swarm,
            "other_configs": others,
        }
        if not args.no_cpu_baseline and world == 1:
            os.environ.setdefault("OMP_NUM_THREADS", "1")
            try:
                kind = _cpu_setup(N)
                n_focal = int(os.environ.get("ABM_BENCH_CPU_FOCAL", 256))
                rate, dt = cpu_rate(n_focal, procs=1)
                line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": kind,
                                        "sample": cpu_sample_text(kind, n_focal, N, 1, dt)}
            except Exception as exc:                                  # noqa: BLE001
                print(f"bench.py: cpu_baseline leg failed: {exc!r}", file=sys.stderr, flush=True)
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "failed: " + repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
