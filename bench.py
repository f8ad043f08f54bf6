#!/usr/bin/env python
"""bench.py -- Llama-3-8B AWQ-INT4 batch-1 decode on B200 (BASELINE.json configs[1]), the reference CPU path beside it.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own AVX CPU path (rank 0 only)

A "step" is one decode token of the synthetic Llama-3-8B (random AWQ-INT4 weights in the reference's QM_CUDA
layout, random-filled fp16 KV cache).  The K timed steps are spread evenly over context lengths 1 -> max_ctx, so
ms_per_step estimates the mean cost per token of a 1 -> 4096 generation.  `value` = tokens/s with token ids already
in HBM (tce_llama_decode); `e2e` = the same through the host entry point (tce_llama_decode_host): token id/position
copied from pinned host memory and the fp32 logits row + greedy token copied back, every step, inside the timed
region.  One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the roofline arithmetic.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "llama3_8b_awq_int4_batch1_decode_tokens_per_s"
UNIT = "tok/s"


def load_traffic():
    """dram__bytes_read + dram__bytes_write per GEMV launch from the committed ncu capture (tools/ncu_launch_summary.py); None if absent."""
    p = Path(__file__).resolve().parent / "profiles" / "roofline_traffic.json"
    try:
        return json.loads(p.read_text()).get("dram_bytes_per_launch")
    except Exception:
        return None


def workload_config(geom, args, world: int, tp: bool) -> dict:
    """The `config` object both arms report: names the workload, no model-architecture keys."""
    return {
        "workload": (f"{geom.name} AWQ-INT4 g128 batch-1 decode, timed steps spread over ctx 1->{args.max_ctx}" if args.ctx < 0
                     else f"{geom.name} AWQ-INT4 g128 batch-1 decode at ctx {args.ctx}"),
        "sequences": 1 if (world == 1 or tp) else world,
        "max_ctx": args.max_ctx,
        "parallelism": (f"tp{world}: one sequence, column/row sharded linears, 2 all-reduces per layer over NVLink peer memory" if tp else
                        (f"{world} independent sequences, one per GPU (no collective)" if world > 1 else "single GPU")),
    }


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ------------------------------------------------------------------------------------------------------------------
# clocks: sample NVML during the timed region (B200_PROFILING.md timing hygiene)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's AVX W4A8 path (oracle/_ref/libtce_ref_avx.so) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_decode(geom, budget_s: float = 20.0):
    """Time Linear_FP_int4::forward's kernel (mat_mul_accelerator_int8_int4_fast_no_offset, QM_x86 g32 format) over
    the linears of ONE decoder layer + lm_head at `geom` shapes with all host threads, extrapolate to a token:
    t_token = L * t_layer + t_lm_head.  Returns dict(value tok/s, cores, kind, sample)."""
    import numpy as np

    from oracle import capi

    cores = os.cpu_count() or 1
    if capi.ref_available("avx"):
        kind = "reference"
        X = capi.ref("avx")
    else:
        kind = "port"
        X = None
    hd = geom.head_dim
    E, F = geom.embed_dim, geom.hidden_dim
    layer_shapes = [(geom.num_heads * hd, E), (geom.num_kv_heads * hd, E), (geom.num_kv_heads * hd, E), (E, geom.num_heads * hd), (F, E), (F, E), (E, F)]
    rng = np.random.default_rng(0)

    def make(oc, ic):
        B = capi.aligned_empty((oc, ic // 2), np.uint8)
        B[:] = rng.integers(0, 256, (oc, ic // 2), dtype=np.uint8)
        S = capi.aligned_empty((oc, ic // 32), np.float32)
        S[:] = (rng.random((oc, ic // 32), dtype=np.float32) + 0.5) * 0.004
        A = capi.aligned_empty((1, ic), np.float32)
        A[:] = rng.standard_normal((1, ic), dtype=np.float32)
        Cc = capi.aligned_empty((1, oc), np.float32)
        xi8 = capi.aligned_empty((ic,), np.int8)
        xs = capi.aligned_empty((ic // 32,), np.float32)
        return A, B, S, Cc, xi8, xs

    def run(t, oc, ic):
        A, B, S, Cc, xi8, xs = t
        if X is not None:
            X.ref_w4a8_avx(A.ctypes.data, B.ctypes.data, S.ctypes.data, Cc.ctypes.data, xi8.ctypes.data, xs.ctypes.data, 1, ic, oc, cores)
        else:  # oracle port of the naive path (scalar, 1 core): only when oracle/_ref could not be built
            capi.naive_mat_mul_int4(np.asarray(A), np.asarray(B), np.asarray(S), 8.0, 32)

    if X is None:
        cores = 1
    mats = [(make(oc, ic), oc, ic) for oc, ic in layer_shapes]
    for t, oc, ic in mats:
        run(t, oc, ic)  # warm-up (also creates the reference's static thread pool with `cores` threads)
    t0 = time.perf_counter()
    reps = 0
    while True:
        for t, oc, ic in mats:
            run(t, oc, ic)
        reps += 1
        if time.perf_counter() - t0 > budget_s * 0.6 or reps >= 20:
            break
    t_layer = (time.perf_counter() - t0) / reps
    lm = make(geom.vocab_size, E)
    run(lm, geom.vocab_size, E)
    t1 = time.perf_counter()
    lm_reps = 0
    while True:
        run(lm, geom.vocab_size, E)
        lm_reps += 1
        if time.perf_counter() - t1 > budget_s * 0.3 or lm_reps >= 10:
            break
    t_lm = (time.perf_counter() - t1) / lm_reps
    t_token = geom.num_layers * t_layer + t_lm
    return {"value": 1.0 / t_token, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"linears of 1 decoder layer x{reps} + lm_head x{lm_reps} at {geom.name} shapes (W4A8 g32, the reference's CPU format), "
                      f"extrapolated t_token = {geom.num_layers}*t_layer + t_lm_head = {t_token * 1e3:.1f} ms; attention/norms not included"}


def run_reference(args):
    rank, _, world = env_rank()
    if rank != 0:
        return
    from tinychatengine_b200.llama import GEOMETRIES

    geom = GEOMETRIES[args.model]
    samples = []
    for _ in range(max(1, args.warmup // 3)):
        cpu_reference_decode(geom, budget_s=4.0)
    t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 3))):
        samples.append(cpu_reference_decode(geom, budget_s=max(4.0, 60.0 / max(1, min(args.steps, 3)))))
    best = max(samples, key=lambda d: d["value"])
    med = sorted(s["value"] for s in samples)[len(samples) // 2]
    cfg = workload_config(geom, args, 1, False)  # the same workload as our arm; the reference has no multi-GPU path: rank 0's host cores
    cfg["reference_path"] = "the reference's AVX W4A8 kernels (oracle/_ref, compiled in place) on all host threads; each step = a bounded sample"
    cfg["sampled_steps"] = len(samples)
    line = {"impl": "reference", "metric": METRIC, "value": med, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / med, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "w4a8 (int8 act x int4 weight, fp32 acc)",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": dict(best, value=med),
            "e2e": {"value": med, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, local_rank, world = env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- tinychatengine_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from tinychatengine_b200 import llama as LL
    from tinychatengine_b200.runtime import Context

    geom = LL.GEOMETRIES[args.model]
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        ctx = Context(local_rank, stream)
        tp = world > 1 and args.parallel == "tp"
        if tp:
            # tensor-parallel decode of ONE sequence: every rank holds a 1/N shard of every matrix (random shards of the
            # right shapes), all-reduce over NVLink peer memory inside the GEMV kernels
            if geom.num_kv_heads % world or geom.num_heads % world:
                raise SystemExit(f"{geom.name}: heads do not split over {world} ranks")
            gl = LL.LlamaGeometry(geom.name, geom.num_layers, geom.num_heads // world, geom.num_kv_heads // world, geom.embed_dim,
                                  geom.hidden_dim // world, geom.vocab_size // world, geom.rms_eps, geom.rope_theta, geom.head_dim)
            model = LL.LlamaModel(ctx, gl, max_ctx=args.max_ctx, seed=1234 + rank, tp_rank=rank, tp_size=world)
            model.tp_connect()
        else:
            gl = geom
            model = LL.LlamaModel(ctx, geom, max_ctx=args.max_ctx, seed=1234 + rank)
        # random-filled KV cache so every context length is "already generated"
        for l in range(geom.num_layers):
            model.kv_cache(l, 0).normal_(0, 0.5)
            model.kv_cache(l, 1).normal_(0, 0.5)
        K, W = args.steps, args.warmup
        gen = torch.Generator(device="cpu")
        gen.manual_seed(99)
        toks = torch.randint(0, geom.vocab_size, (K + W,), generator=gen).tolist()
        # timed positions spread over 1 -> max_ctx (ctx length = pos); warm-up at mid context
        pos_list = [args.max_ctx // 2] * W + [min(args.max_ctx - 1, int(round(i * (args.max_ctx - 1) / max(1, K - 1)))) for i in range(K)]
        if args.ctx >= 0:
            pos_list = [args.ctx] * (W + K)
        tokpos_all = torch.tensor(list(zip(toks, pos_list)), dtype=torch.int32, device=dev)
        tokpos = torch.zeros(2, dtype=torch.int32, device=dev)
        logits_pinned = torch.empty(gl.vocab_size, dtype=torch.float32).pin_memory()

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        # ---- value: token ids resident in HBM ----
        for i in range(W):
            tokpos.copy_(tokpos_all[i])
            model.decode(tokpos)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            e0.record(stream)
            for i in range(W, W + K):
                tokpos.copy_(tokpos_all[i])
                model.decode(tokpos)
            e1.record(stream)
            barrier()
        ms_dev = e0.elapsed_time(e1)
        # ---- e2e: host entry point, H2D + D2H inside ----
        for i in range(W):
            model.decode_host(toks[i], pos_list[i], logits_pinned)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for i in range(W, W + K):
            model.decode_host(toks[i], pos_list[i], logits_pinned)
        e3.record(stream)
        barrier()
        ms_e2e = e2.elapsed_time(e3)
        # ---- dominant kernel: the W4A16 GEMV launches of one step, timed alone with events ----
        if tp:
            n_gemv, ms_gemv_step = 4 * geom.num_layers + 1, None  # the sharded GEMVs wait for their peers: not timed alone
        else:
            n_gemv = ctx.L.tce_llama_enqueue_gemvs(model.h)  # eager once: modules loaded, attributes set
            barrier()
            reps = 20
            # replayed from a CUDA graph like the real step (plain stream launches would add ~2 us of launch gap per kernel)
            gemv_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gemv_graph, stream=stream):
                for _ in range(reps):
                    ctx.L.tce_llama_enqueue_gemvs(model.h)
            with torch.cuda.stream(stream):
                gemv_graph.replay()
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(stream)
                gemv_graph.replay()
                g1.record(stream)
            barrier()
            ms_gemv_step = g0.elapsed_time(g1) / reps

    times = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = times.tolist()
    if rank == 0:
        peak, peak_src = load_peaks()
        wbytes = LL.weight_bytes_per_token(geom)
        mean_ctx = sum(pos_list[W:]) / K
        kvbytes = LL.kv_bytes_per_token(geom, int(mean_ctx))
        seqs = 1 if tp else world
        tok_s = seqs * K / (ms_dev * 1e-3)
        e2e_tok_s = seqs * K / (ms_e2e * 1e-3)
        gemv_gbs = None if tp else wbytes / (ms_gemv_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": tok_s, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K,
            "higher_is_better": True, "scaling": "strong" if tp else "weak", "vs_baseline": None,
            "dtype": "w4a16 (int4 weights; activations fp16 -> 15-bit block fixed point; int32 accumulate)",
            "data": "synthetic",
            "config": dict(workload_config(geom, args, world, tp), mean_ctx=mean_ctx,
                           l2="inputs larger than L2: 3.9 GB of weights stream per step vs 126 MB L2", pdl=bool(int(os.environ.get("TCE_USE_PDL", "1")))),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_tok_s, "unit": UNIT, "h2d_bytes_per_step": 12, "d2h_bytes_per_step": gl.vocab_size * 4 + 4},
            "gpu_launches": K * model.kernels_per_step,
            "roofline": {"bound": "hbm", "kernel": f"w4a16_gemv_kernel<1> ({n_gemv} launches/step: 4 per layer + lm_head)",
                         "achieved": gemv_gbs, "peak": peak, "unit": "GB/s", "frac": None if tp else gemv_gbs / peak, "traffic": load_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": wbytes / n_gemv, "avg_launch_us": None if tp else ms_gemv_step * 1e3 / n_gemv,
                         "step": {"bytes_per_token": wbytes + kvbytes, "achieved": (wbytes + kvbytes) / (ms_dev / K * 1e-3) / 1e9,
                                  "frac": (wbytes + kvbytes) / (ms_dev / K * 1e-3) / 1e9 / peak}},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_reference_decode(geom, budget_s=args.cpu_budget)
            except Exception as ex:  # the baseline must never hide the GPU number
                line["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(line), flush=True)
    model.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--max-ctx", type=int, default=4096)
    ap.add_argument("--ctx", type=int, default=-1, help="fixed context length for every step (default: sweep 1 -> max_ctx)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # N > 1 default = one independent batch-1 sequence per GPU (no data-path collective, weak scaling): the throughput configuration.
    # --parallel tp = tensor-parallel decode of ONE sequence over NVLink peer memory (BASELINE config 5, strong scaling): implemented
    # and parity-tested at 2 GPUs, measured 464 / 427 tok/s at 2 / 4 GPUs vs 566 on one (profiles/README.md) -- batch-1 latency is
    # bound by per-launch cost, not bandwidth, so sharding the weights does not pay yet; not verified at 8 GPUs this round.
    ap.add_argument("--parallel", default="replicas", choices=["tp", "replicas"], help="N>1: one sequence per GPU (default), or tensor-parallel decode of one sequence")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
