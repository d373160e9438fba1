#!/usr/bin/env python
"""bench.py -- pair.sites/s through the Phyloformer forward on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 200x1000] [--precision bf16x3]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores

One "step" = one full forward (pair embedding -> 6 axial blocks -> distance head) of one
synthetic MSA batch.  Default workload: BASELINE config 3, one 200-taxon x 1000-site MSA
(19 900 pairs, 19.9 M pair.sites, 5.1 GB of fp32 activations -- far larger than L2, so no
flush is needed between steps).  With N > 1 the pair axis of the SAME MSA is sharded over the
ranks (one all-reduce of (L,72) floats per block + one gather of the distances), i.e. strong
scaling.  Prints ONE JSON line (rank 0).

  value      device-timed throughput, MSA codes already resident in HBM (forward_idx)
  e2e        the same metric through the reference-facing call: host (pinned) fp32 one-hot
             (B,22,L,n) -> .cuda() -> model(x) -> .cpu(), copies inside the timed region
  roofline   the dominant kernel (column-apply + LN + FFN + residual), timed live with CUDA
             events around each launch (pf_profile_*), against MEASURED_PEAKS.json; roofline.kernels
             holds the same for the row / column-summary kernels against the HBM peak
  cpu_baseline  the unmodified reference module (baseline/_ref, staged by tools/stage_ref.sh; the CPU
             oracle port when it is absent) on a bounded sample of the same workload, on this box's
             host cores
  checks     parity of what was just timed: against the full-size fp64 oracle fixture, and at N > 1
             against the unsharded forward and across ranks (bit equality)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pair_sites_per_s"
UNIT = "pair*sites/s"
FLOP_PER_TOKEN_FFN = 65536.0          # 2*(64*256 + 256*64)            SURVEY 8(a)
BYTES_PER_TOKEN_FFN = 512.0           # one fp32 read + one fp32 write of the 64-vector
BYTES_PER_TOKEN_FORWARD = 3072.0      # 12 activation passes           SURVEY 8(d)
WORKLOADS = {                         # name -> (B, n, L)
    "200x1000": (1, 200, 1000),       # BASELINE config 3 (north-star target shape)
    "50x500": (1, 50, 500),           # config 2
    "100x500": (1, 100, 500),         # the MSAs/s shape of the metric
    "256x20x200": (256, 20, 200),     # config 4 (batched small MSAs)
    "500x500": (1, 500, 500),         # config 5 (needs >= 2 GPUs worth of HBM headroom: 16 GB, fits one)
}


def workload_desc(name):
    B, n, L = WORKLOADS[name]
    return f"PF {name}: B={B} MSAs of {n} taxa x {L} sites, {n * (n - 1) // 2} pairs each, pf.ckpt weights"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="200x1000", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("PF_PRECISION", "bf16x3"),
                    choices=["fp32", "bf16x3", "bf16", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "p2p"],
                    help="cross-shard sum of the column summaries: library kernels over NVLink peer memory, or NCCL")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def load_weights_sd():
    import torch
    ck = torch.load(os.path.join(ROOT, "tests", "golden", "ckpt_pf.pt"), map_location="cpu")
    return ck


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": max(pw)}
        return out


# --------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def load_reference_forward():
    """The CPU comparator: (callable x -> distances, kind).
    kind "reference": the UNMODIFIED reference module (phyloformer/model.py + attention.py, staged byte
    for byte under baseline/_ref/phyloformer_ref/ by tools/stage_ref.sh -- git-ignored, it travels to
    the GPU box with the snapshot), built and loaded exactly like reference infer_alns.py:71-86.
    kind "port": the CPU oracle (oracle/pf_oracle.py), used only when the staged files are absent."""
    import torch
    ck = load_weights_sd()
    if os.path.exists(os.path.join(REF_DIR, "phyloformer_ref", "model.py")):
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from phyloformer_ref.model import Phyloformer as RefPhyloformer
        m = RefPhyloformer(**ck["hyper_parameters"])
        m.load_state_dict({k.replace("model.", ""): v for k, v in ck["state_dict"].items() if k != "model.seq2pair"},
                          strict=False)
        m.eval()
        return (lambda x: m(x)), "reference", m
    from oracle import pf_oracle
    w = pf_oracle.strip_prefix(ck["state_dict"])
    return (lambda x: pf_oracle.forward(w, x, torch.float32)), "port", None


def cpu_ref_rate(n, L, threads, repeats=1):
    """pair.sites/s of the CPU comparator (fp32, torch CPU ops) on one n x L MSA."""
    import torch
    from oracle import pf_oracle
    torch.set_num_threads(threads)
    fwd, kind, _ = load_reference_forward()
    idx = pf_oracle.synth_msa(n, L, seed=1337, kind="uniform")
    x = pf_oracle.msa_to_onehot(idx)
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            fwd(x)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return (n * (n - 1) // 2) * L / best, best, kind


def pick_cpu_sample(threads, budget_s):
    """Size a sample of the workload (fewer taxa/sites, same per-token work) to ~budget_s."""
    if os.environ.get("PF_BENCH_CPU_SAMPLE"):               # "taxa x sites", e.g. 8x32 (tests)
        n, L = (int(v) for v in os.environ["PF_BENCH_CPU_SAMPLE"].lower().split("x"))
        return n, L
    rate, _, _ = cpu_ref_rate(16, 128, threads)             # quick probe (~0.2 s)
    tokens = max(rate * budget_s, 2e4)
    for n, L in ((100, 500), (64, 500), (48, 400), (40, 250), (30, 200), (20, 200), (16, 128)):
        if (n * (n - 1) // 2) * L <= tokens:
            return n, L
    return 16, 128


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (the
    unmodified module from baseline/_ref when staged, else the oracle port), same metric/config keys.
    The reference cannot hold the full workloads in host memory (~3.5 KB per pair.site: 70 GB at
    200 x 1000) and refuses n > 200 (model.py:24-28); each step is a bounded sample of the workload --
    its cost is exactly linear in pairs x sites."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    per_step = max(2.0, min(10.0, 200.0 / max(1, args.steps + args.warmup)))
    sn, sL = pick_cpu_sample(threads, per_step)
    from oracle import pf_oracle
    torch.set_num_threads(threads)
    fwd, kind, _ = load_reference_forward()
    x = pf_oracle.msa_to_onehot(pf_oracle.synth_msa(sn, sL, seed=1337, kind="uniform"))
    with torch.no_grad():
        for _ in range(args.warmup):
            fwd(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fwd(x)
        dt = time.perf_counter() - t0
    tokens = (sn * (sn - 1) // 2) * sL
    val = tokens * args.steps / dt
    what = ("unmodified reference phyloformer.model.Phyloformer.forward (baseline/_ref)" if kind == "reference"
            else "CPU oracle port of the reference graph")
    sample = f"{sn} taxa x {sL} sites ({tokens} pair*sites per step), fp32, torch CPU ops, {what}"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload), "tokens_per_step": tokens, "precision": "fp32",
                   "parallelism": f"host CPU, {threads} threads (rank 0 only)",
                   "sample": f"each step is a bounded sample of the workload: {sample}"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from oracle import pf_oracle  # synthetic generator + cpu_baseline leg only
    from phyloformer_b200.model import Phyloformer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, n, L = WORKLOADS[args.workload]
    P = n * (n - 1) // 2
    tokens = B * P * L

    ck = load_weights_sd()
    model = Phyloformer(**ck["hyper_parameters"], precision=args.precision)
    model.load_state_dict({k.replace("model.", ""): v for k, v in ck["state_dict"].items()
                           if k != "model.seq2pair"}, strict=False)
    model = model.to(dev).eval()
    batch_sharded = world > 1 and B >= world
    if world > 1 and not batch_sharded:
        model.shard_pairs(exchange=args.exchange)
        model.check_peer_errors = False      # keep the timed forwards asynchronous; the flag is read after the timed regions
    idx_host = pf_oracle.synth_msa(n, L, seed=1337 + n, kind="tree", B=B)
    if batch_sharded:  # independent MSAs: replicas, no collective
        from phyloformer_b200 import sharding
        lo, hi = sharding.batch_range(B, rank, world)
        idx_host = idx_host[lo:hi].contiguous()
    idx = idx_host.to(dev)
    x_host = pf_oracle.msa_to_onehot(idx_host).pin_memory()      # the reference-facing input

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            d = model.forward_idx(idx, squeeze=False)
        barrier()
        # ---- timed region 1: device-resident inputs --------------------------------------
        model.profile_enable(True)
        model.profile_read()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = 0
        barrier()
        e0.record()
        for _ in range(args.steps):
            d = model.forward_idx(idx, squeeze=False)
            launches += model.last_launches
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None
        prof = model.profile_read()
        model.profile_enable(False)
        # ---- timed region 2: end to end through forward(x) with host buffers ---------------
        # Every step copies its input host -> device (pinned memory) and reads its result back; as in the CLI's pipeline
        # (infer_alns.run_pipeline) the copy of step k+1 is issued on a copy stream while step k computes (two device
        # input buffers), and the result read of step k is what the host waits for.
        for _ in range(2):
            model(x_host.to(dev, non_blocking=True)).cpu()
        copy_stream = torch.cuda.Stream(device=dev)
        x_dev = [torch.empty(x_host.shape, dtype=x_host.dtype, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        main_stream = torch.cuda.current_stream(dev)

        def prefetch(slot):     # the forward that last read this buffer (step k-1) has completed: its result was read back
            with torch.cuda.stream(copy_stream):
                x_dev[slot].copy_(x_host, non_blocking=True)
                ready[slot].record(copy_stream)

        barrier()
        e0.record()
        prefetch(0)
        for k in range(args.steps):
            cur = k & 1
            main_stream.wait_event(ready[cur])
            out = model(x_dev[cur])
            if k + 1 < args.steps:
                prefetch(cur ^ 1)
            out_host = out.cpu()
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        # ---- correctness of what was just timed (outside the timed regions) --------------------
        checks = {}
        model.check_device_error()      # a bounded device-side wait that timed out invalidates the run: fail loudly
        fx = os.path.join(ROOT, "tests", "golden", f"oracle_fullsize_{n}x{L}.npz")
        if B == 1 and os.path.exists(fx):     # every distance against the pair-chunked fp64 oracle (tests/golden/make_fullsize.py)
            import numpy as np
            ref = torch.from_numpy(np.load(fx)["dist"]).to(dev)
            checks["parity_vs_oracle_fixture_max_rel"] = float(((d[0].double() - ref).abs() / ref).max())
        if world > 1 and not batch_sharded:
            # (a) all ranks hold bit-identical gathered results; (b) rank 0 recomputes the forward unsharded
            mine = d.contiguous().view(torch.int32)
            allv = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine)
            checks["cross_rank_bit_equal"] = bool(all(torch.equal(allv[0], v) for v in allv))
            model.check_device_error()
            if rank == 0:
                model.unshard()
                d_full = model.forward_idx(idx, squeeze=False)
                torch.cuda.synchronize()
                checks["parity_vs_unsharded_max_rel"] = float(((d.double() - d_full.double()).abs() / d_full.double().abs()).max())
                del d_full
            barrier()
            model.shard_pairs(exchange=args.exchange)
        # ---- MSAs/s at 100 x 500 (second half of BASELINE.json's metric), at every N ---------
        # replicas: every rank runs its own 100 x 500 alignment (no collective): N x K alignments / max-over-ranks time
        msas = {}
        saved_shard = model._shard
        model.unshard()
        i2 = pf_oracle.synth_msa(100, 500, seed=1437 + rank).to(dev)
        for _ in range(3):
            model.forward_idx(i2)
        barrier()
        e0.record()
        for _ in range(10):
            model.forward_idx(i2)
        e1.record()
        barrier()
        msas["replicas"] = world * 10 / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
        if world > 1:     # the same alignment pair-sharded over all ranks (latency of one alignment)
            model.shard_pairs(exchange=args.exchange)
            i3 = pf_oracle.synth_msa(100, 500, seed=1437).to(dev)
            for _ in range(3):
                model.forward_idx(i3)
            barrier()
            e0.record()
            for _ in range(10):
                model.forward_idx(i3)
            e1.record()
            barrier()
            msas["pair_sharded"] = 10 / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
        model._shard = saved_shard
        msas_per_s = msas["replicas"]
        # ---- the reference module itself, eager PyTorch on this GPU (100 x 500; seq2pair device shim of
        #      the reference leaves its seq2pair matrix on the CPU, model.py:136,200-201) ----
        ref_gpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            try:
                _, kind, rm = load_reference_forward()
                if kind == "reference":
                    rm = rm.to(dev)
                    rm._set_seq2pair(100)
                    rm.seq2pair = rm.seq2pair.to(dev)
                    xr = pf_oracle.msa_to_onehot(pf_oracle.synth_msa(100, 500, seed=1437)).to(dev)
                    for _ in range(2):
                        dr = rm(xr)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(5):
                        dr = rm(xr)
                    e1.record()
                    torch.cuda.synchronize()
                    ms_ref = e0.elapsed_time(e1) / 5
                    ours = model.forward_idx(pf_oracle.synth_msa(100, 500, seed=1437).to(dev))
                    ref_gpu = {"what": "unmodified reference module, eager PyTorch fp32 on this B200 (seq2pair moved to the device)",
                               "workload": "100 taxa x 500 sites", "ms_per_forward": ms_ref, "msas_per_s": 1e3 / ms_ref,
                               "pair_sites_per_s": 4950 * 500 / (ms_ref * 1e-3),
                               "native_vs_this_max_rel": float(((ours.double() - dr.double()).abs() / dr.double().abs()).max())}
                    del rm, xr, dr
                    torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001  (a comparator, not the product)
                ref_gpu = {"unavailable": repr(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    ms_step = ms_total / args.steps
    value = tokens / (ms_step * 1e-3)
    local_tokens = tokens / world
    # dominant kernel: column-apply + LN + FFN + residual, one launch per block
    ffn_ms, ffn_n = prof["ffn"]
    ffn_avg = ffn_ms / max(ffn_n, 1)
    kernels = {}
    for k, (ms, cnt) in prof.items():
        if cnt:
            kernels[k] = {"ms_per_step": ms / args.steps, "launches_per_step": cnt / args.steps}
    share = ffn_ms / args.steps / ms_step if ms_step > 0 else None
    if args.precision == "fp32":
        gbs = local_tokens * BYTES_PER_TOKEN_FFN / (ffn_avg * 1e-3) / 1e9
        roofline = {"kernel": "k_colapply_ffn_fp32", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                    "note": "fp32 FFMA mode is CUDA-core bound; HBM figure given for reference"}
    else:
        issued = {"bf16x3": 3.0, "fp16": 2.0}.get(args.precision, 1.0)
        tf = local_tokens * FLOP_PER_TOKEN_FFN / (ffn_avg * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        roofline = {"kernel": "k_colapply_ffn_ws" if os.environ.get("PF_FFN_IMPL", "ws") != "tc" else "k_colapply_ffn_tc", "bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s",
                    "frac": tf / peak, "traffic": None, "achieved_issued": tf * issued, "frac_issued": tf * issued / peak,
                    "hbm_gbs": local_tokens * BYTES_PER_TOKEN_FFN / (ffn_avg * 1e-3) / 1e9,
                    "note": f"useful FLOPs = 65536/token; {int(issued)} 16-bit MMA passes issued per product"}
    # the streaming kernels against the HBM roofline: algorithmic bytes per launch / live launch time
    nb = 6
    hbm_alg = {"row": local_tokens * (512.0 * (nb - 1) + 256.0) / nb,   # read + write in place; block 0 only writes (x0 comes from the MSA)
               "colsum": local_tokens * 256.0,                          # one read pass
               "ffn": local_tokens * (512.0 * nb - (256.0 if prof.get("head", (0, 0))[0] < 0.2 * args.steps else 0.0)) / nb}
    for k, alg in hbm_alg.items():
        if k in kernels and kernels[k]["launches_per_step"] > 0:
            per_launch_ms = kernels[k]["ms_per_step"] / kernels[k]["launches_per_step"]
            gbs = alg / (per_launch_ms * 1e-3) / 1e9
            kernels[k].update({"avg_launch_ms": per_launch_ms, "algorithmic_bytes_per_launch": alg, "hbm_gbs": gbs,
                               "hbm_frac": gbs / peaks["hbm_gbs"]})
    # DRAM traffic of the dominant kernel from the committed ncu capture (same workload, 1 GPU)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if tr["workload"] == args.workload and world == tr["n_gpus"] and args.precision != "fp32":
            roofline["traffic"] = tr["dram_bytes_per_launch"]["k_colapply_ffn_ws"]
            roofline["traffic_algorithmic"] = tr["algorithmic_bytes_per_launch"]["k_colapply_ffn_ws"]
            for cls, kname in (("row", "k_row_attn_ws"), ("colsum", "k_col_partial_ws"), ("ffn", "k_colapply_ffn_ws")):
                if cls in kernels:      # ncu --set full capture of one launch of the class's kernel (profiles/r02_ncu_summary.txt)
                    kernels[cls]["traffic"] = tr["dram_bytes_per_launch"][kname]
                    kernels[cls]["ncu_kernel"] = kname
    except Exception:  # noqa: BLE001
        pass
    roofline.update({"peak_source": peaks["source"] + " (sustained)" if roofline["bound"] == "tensor" else peaks["source"],
                     "avg_launch_ms": ffn_avg, "share_of_step": share, "kernels": kernels,
                     "forward_hbm_frac_at_3072B": value * BYTES_PER_TOKEN_FORWARD / 1e9 / peaks["hbm_gbs"] / world})

    h2d = x_host.numel() * 4
    d2h = out_host.numel() * out_host.element_size()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if batch_sharded else "strong", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "f32 (FFN: bf16x3 tcgen05, f32 accumulate)", "bf16": "f32 (FFN: bf16 tcgen05)",
                  "fp16": "f32 (FFN: fp16 activations x fp16 hi+lo weights, tcgen05, f32 accumulate)"}[args.precision],
        "data": "synthetic",
        "config": {"workload": workload_desc(args.workload),
                   "tokens_per_step": tokens, "precision": args.precision,
                   "parallelism": ("1 GPU" if world == 1 else (f"batch-sharded x{world} (replicas)" if batch_sharded
                                                               else f"pair-sharded x{world}, sum of (L,72) fp32 per block over "
                                                                    + ("NVLink peer memory (own kernels)" if getattr(model, "_peer", None) else "NCCL all-reduce"))),
                   "l2": "activations (%.2f GB per rank) exceed L2; no flush needed" % (local_tokens * 256 / 1e9)},
        "clocks": clocks,
        "e2e": {"value": tokens / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                "api": "Phyloformer.forward(x: (B,22,L,n) fp32 one-hot copied from pinned host memory every step; the copy of step k+1 overlaps step k).cpu()"},
        "gpu_launches": launches,
        "roofline": roofline,
    }
    line["msas_per_s_100x500"] = msas_per_s
    line["msas_100x500"] = {"replicas_msas_per_s": msas["replicas"], "pair_sharded_msas_per_s": msas.get("pair_sharded"),
                            "note": "replicas: every rank runs its own alignment; pair_sharded: one alignment over all ranks"}
    line["checks"] = checks
    if ref_gpu is not None:
        line["reference_gpu_eager"] = ref_gpu
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sn, sL = pick_cpu_sample(threads, 15.0)
        rate, secs, kind = cpu_ref_rate(sn, sL, threads)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                                "sample": f"{sn} taxa x {sL} sites, one forward, {secs:.1f} s, fp32 torch CPU ops, "
                                          + ("unmodified reference module (baseline/_ref)" if kind == "reference" else "oracle port")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
