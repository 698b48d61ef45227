#!/usr/bin/env python
"""Benchmark of the Multi-ATGCN train step (BASELINE.json metric: train samples/s).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the reference

One "step" = one pass of ``TrafficStateExecutor._train_epoch``'s loop body over one batch of
synthetic input of the named shape: zero_grad -> calculate_loss (forward) -> backward ->
[gradient all-reduce when N > 1] -> clip_grad_norm_(5) -> Adam step (the last two fused over one flat
bucket, ``train.FusedClipAdam``).

* ``value``  : samples/s with the step's inputs already resident in HBM (CUDA-event timed,
               max over ranks).
* ``e2e``    : the same metric through the public model API with HOST buffers: every step copies
               its batch from pinned host memory to the device (on a copy stream, overlapping the
               previous step) and reads the loss back on the host, all inside the timed region.
* ``e2e_device_windows``: the same step with the batch gathered on the device from a series
               resident in HBM (``train.DeviceWindowBank``); only label-start indices are uploaded.
* ``roofline``: the dominant kernel - the persistent reverse-time recurrence kernel in the default bf16 mode (the forward
               one beside it), the support-propagation GEMM in the other modes - timed live with CUDA events on the
               launching stream.
* ``cpu_baseline``: the oracle (a port of the reference's CPU path) timed on this host's cores
               on a bounded sample of the same workload.
* ``strong_scaling``: BASELINE config 4 (N=883, GLOBAL batch 256 sharded over the ranks): the one scaling target the
               north star states (>= 7x at 8 GPUs); a few steps after the main measurement, at every --gpus N.
* ``step_roofline``: SURVEY 8d's work model of the whole step against the measured peaks.
Multi-GPU: launched by torchrun, one rank per GPU, weak scaling (fixed per-GPU batch) for ``value``.
"""
from __future__ import annotations

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

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "multi_atgcn_train_samples_per_s"
UNIT = "samples/s"
DEFAULT_WORKLOAD = "baltimore_multi"  # N=403, T=24 -> 24, batch 64 per GPU: the shape the metric is quoted on


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [c.strip() for c in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = sorted(sm)[len(sm) // 2:]  # upper half ~ samples taken under load
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def _dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
    return world, rank, local


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle on host cores
# ------------------------------------------------------------------------------------------------
def _oracle_step_time(cfg, df, params, n_nodes, t_out, batch, seed=0):
    from multistgraph_b200.synthetic import make_batch
    from oracle.matgcn_oracle import OracleModel

    model = OracleModel(cfg, df, params)
    b = make_batch(n_nodes, batch, t_out, seed=seed)
    t0 = time.perf_counter()
    loss = model.calculate_loss(b)
    loss.backward()
    return time.perf_counter() - t0


def cpu_reference_run(workload_name: str, steps: int, warmup: int, budget_s: float):
    """Times the CPU restatement of the reference's train step (forward + backward of
    calculate_loss; the reference has no GPU-free fast path) on all host threads, on the
    largest batch <= the workload's that fits ``budget_s``."""
    from multistgraph_b200.model import MultiATGCN
    from multistgraph_b200.synthetic import WORKLOADS, workload

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = WORKLOADS[workload_name]
    cfg, df, _ = workload(workload_name, seed=0, batch=1)
    torch.manual_seed(0)
    params = {k: v.detach().clone() for k, v in MultiATGCN(dict(cfg), df).state_dict().items()}
    t1 = _oracle_step_time(cfg, df, params, w["N"], w["T_out"], 1)           # also warms the thread pool
    t2 = _oracle_step_time(cfg, df, params, w["N"], w["T_out"], 2)
    slope = max(t2 - t1, 0.02 * t1)
    fixed = max(t1 - slope, 0.0)
    k_eff = max(1, min(steps, 3))
    w_eff = 1 if warmup > 0 else 0
    spent = t1 + t2
    batch = 1
    for cand in (w["B"], w["B"] // 2, w["B"] // 4, w["B"] // 8, 4, 2, 1):
        if cand < 1:
            continue
        if (k_eff + w_eff) * (fixed + slope * cand) <= max(budget_s - spent, 1.0):
            batch = cand
            break
    for _ in range(w_eff):
        _oracle_step_time(cfg, df, params, w["N"], w["T_out"], batch)
    times = [_oracle_step_time(cfg, df, params, w["N"], w["T_out"], batch, seed=i) for i in range(k_eff)]
    dt = sum(times) / len(times)
    return {"value": batch / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d timed step(s) of forward+backward at batch %d of %d (N=%d), %.1f s/step, %d torch threads"
                      % (k_eff, batch, w["B"], w["N"], dt, cores),
            "steps": k_eff, "warmup": w_eff, "ms_per_step": dt * 1e3, "batch": batch}


def run_reference_arm(args):
    world, rank, _ = _dist_setup(args.gpus)
    if rank != 0:
        return
    r = cpu_reference_run(args.workload, args.steps, args.warmup, budget_s=args.cpu_budget if args.cpu_budget else 200.0)
    from multistgraph_b200.synthetic import WORKLOADS
    w = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": _workload_desc(args.workload, w, r["batch"]), "device": "host CPU (no GPU code in the reference)"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _workload_desc(name, w, batch):
    return ("%s: N=%d nodes, batch %d per GPU, 24 steps in -> %d out, adjtype=%s adpadj=%s K=5 supports, "
            "embed_dim=%d, hidden 64, 2 layers" % (name, w["N"], batch, w["T_out"], w["adjtype"], w["adpadj"], w["D"]))


# ------------------------------------------------------------------------------------------------
# dominant-kernel roofline, measured live
# ------------------------------------------------------------------------------------------------
def propagation_roofline(n_nodes, batch, hidden, kp, ldm, device, peaks, flags, iters=20):
    """Times the support-propagation kernel at the step's shape ([Kp*N, N] x [N, B*H]) with CUDA
    events on the launching stream; L2 is flushed between launches by rewriting a 256 MB buffer."""
    from multistgraph_b200 import _cabi

    lib = _cabi.lib()
    cols = batch * hidden
    M = torch.randn(kp, n_nodes, ldm, device=device) * 0.05
    X = torch.randn(n_nodes, cols, device=device)
    P = torch.empty(kp, n_nodes, cols, device=device)
    flush = torch.empty(64 * 1024 * 1024, device=device, dtype=torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    bf16 = bool(flags & _cabi.FLAG_BF16)
    if bf16:
        M16, X16 = M.bfloat16(), X.bfloat16()
        P16 = torch.empty(kp, n_nodes, cols, device=device, dtype=torch.bfloat16)

        def launch():   # the launch exactly as a bf16-mode step issues it: bf16 operands in, bf16 twin of the result out
            _cabi.check(lib.matgcn_propagate_fwd_bf16_twin(M16.data_ptr(), kp, n_nodes, ldm, X16.data_ptr(), cols, P16.data_ptr(), st),
                        "propagate_bf16_twin")
    else:
        def launch():
            _cabi.check(lib.matgcn_propagate_fwd(M.data_ptr(), kp, n_nodes, ldm, X.data_ptr(), cols, P.data_ptr(), flags, st),
                        "propagate")
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
    ms = total / iters
    flops = 2.0 * kp * n_nodes * n_nodes * cols
    achieved = flops / (ms * 1e-3) / 1e12
    peak = peaks["bf16_tflops"]
    traffic = None
    prof = os.path.join(ROOT, "profiles", "r1_prop_kernel_ncu_bf16_twin.json" if bf16 else "r1_prop_kernel_ncu.json")
    if flags and os.path.exists(prof):
        with open(prof) as f:
            pj = json.load(f)
        if pj["shape"] == {"Kp": kp, "N": n_nodes, "cols": cols}:  # the ncu --set full capture of this very shape
            traffic = pj["dram_bytes_read"] + pj["dram_bytes_write"]
    kern = ("gemm_tc_kernel<128,A_KC,B_NC,BF16,EpiPlain> (support propagation, tcgen05 kind::f16 bf16 + TMA)" if bf16
            else "gemm_tc_kernel<128,A_KC,B_NC,TF32,EpiPlain> (support propagation, tcgen05 kind::tf32 + TMA)" if flags
            else "gemm_kernel<CfgBig,A_KC,B_NC,EpiPlain> (support propagation, fp32 FFMA)")
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": kern, "launch_ms": ms, "flops_per_launch": flops,
            "algorithmic_bytes_per_launch": (2.0 if bf16 else 4.0) * (kp * n_nodes * ldm + n_nodes * cols)
                                            + (2.0 if bf16 else 4.0) * kp * n_nodes * cols,
            "peak_source": peaks["source"] + " bf16 dense burst",
            "note": "peak is the measured bf16 cuBLAS figure; TF32 tensor-core peak is half of it, the fp32 FFMA "
                    "kernel of exact mode cannot approach either"}


def recurrence_roofline(model, batches, opt, train_step, lib, n_nodes, batch, k_supports, n_adp, t_steps, peaks, steps=3):
    """The dominant kernels of the bf16 path are the two persistent recurrence kernels (csrc/rec_fwd.cuh, rec_bwd.cuh: one
    cooperative launch per layer and direction).  Their launches are timed with CUDA events on the launching stream inside
    real train steps (matgcn_rec_timing); the roofline is stated for the larger one, the reverse-time kernel.
    Algorithmic work per launch (DESIGN.md section 4, H = 64): FLOPs = T [2 * 2 Kp N^2 B H + 2 N B K H 3H + 2 N B H 3H];
    HBM bytes = T [saved activations read + pre-activation gradients written (as stored: fp32, bf16 twins) + the per-node
    weight blocks once (bf16)].  By SURVEY 8d's rule the bound is the larger of the two times: HBM."""
    import ctypes

    H, kp = 64, k_supports - 1
    u = n_nodes * batch * H
    flops = t_steps * (2 * 2.0 * kp * n_nodes * n_nodes * batch * H + 2.0 * n_nodes * batch * k_supports * H * 3 * H + 2.0 * n_nodes * batch * H * 3 * H)
    w_bytes = 2.0 * n_nodes * k_supports * H * 3 * H
    # (the fp32 copy of the main-cell pre-activation gradients DG is written by the first layer only - the inner layers' consumers
    # all read the bf16 twin - and the launches of all layers are averaged below)
    dg32_share = 1.0 / max(1, int(getattr(model, "num_layers", 2)))
    bytes_bwd = t_steps * (9 * u * 4.0 + 3 * u * 4.0 * (1.0 + dg32_share) + 3 * u * 2.0 + 2 * n_adp * u * 2.0 + w_bytes)
    bytes_fwd = t_steps * (7 * u * 4.0 + 10 * u * 4.0 + 2 * k_supports * u * 2.0 + w_bytes)
    lib.matgcn_rec_timing(1)
    for i in range(steps):
        train_step(model, batches[i % len(batches)], opt)
    torch.cuda.synchronize()
    f_ms, b_ms = ctypes.c_double(), ctypes.c_double()
    f_n, b_n = ctypes.c_int(), ctypes.c_int()
    lib.matgcn_rec_timing_read(ctypes.byref(f_ms), ctypes.byref(f_n), ctypes.byref(b_ms), ctypes.byref(b_n))
    lib.matgcn_rec_timing(0)
    if f_n.value == 0 or b_n.value == 0:
        return None
    fwd_ms, bwd_ms = f_ms.value / f_n.value, b_ms.value / b_n.value
    traffic = None
    prof = os.path.join(ROOT, "profiles", "r2m_rec_bwd_ncu.json")
    if os.path.exists(prof):
        with open(prof) as f:
            pj = json.load(f)
        if pj.get("shape") == {"N": n_nodes, "B": batch, "K": k_supports, "T": t_steps}:
            traffic = pj["dram_bytes_read"] + pj["dram_bytes_write"]
    achieved = bytes_bwd / (bwd_ms * 1e-3) / 1e9
    head = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"]}
    if flops / (peaks["bf16_tflops_sustained"] * 1e12) > bytes_bwd / (peaks["hbm_gbs"] * 1e9):
        # large graphs (the N^2 term of the dense phases): t_flop > t_hbm, the tensor roofline bounds the launch
        tf = flops / (bwd_ms * 1e-3) / 1e12
        head = {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": tf / peaks["bf16_tflops_sustained"], "hbm": head}
    return {**head,
            "traffic": traffic,
            "kernel": "rec_bwd_kernel (persistent reverse-time recurrence of one layer: %d steps x 4 phases, TMA + tcgen05 + TMEM, "
                      "grid barriers)" % t_steps,
            "launch_ms": bwd_ms, "launches_timed": b_n.value, "algorithmic_bytes_per_launch": bytes_bwd, "flops_per_launch": flops,
            "tensor": {"achieved": flops / (bwd_ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                       "frac": flops / (bwd_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"]},
            "rec_fwd_kernel": {"launch_ms": fwd_ms, "launches_timed": f_n.value, "algorithmic_bytes_per_launch": bytes_fwd,
                               "flops_per_launch": flops, "hbm_frac": bytes_fwd / (fwd_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                               "tensor_frac": flops / (fwd_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"]},
            "peak_source": peaks["source"] + " HBM copy bandwidth; bf16 dense sustained for the tensor view",
            "note": "timed inside real train steps with CUDA events on the launching stream (both layers' launches averaged); "
                    "the bound is the larger of t_hbm and t_flop for the launch (SURVEY 8d: take the larger time)"}


# ------------------------------------------------------------------------------------------------
# this repository's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    from multistgraph_b200 import _cabi
    from multistgraph_b200.dp import broadcast_parameters
    from multistgraph_b200.model import MultiATGCN
    from multistgraph_b200.synthetic import WORKLOADS, make_batch, workload
    from multistgraph_b200.train import DeviceWindowBank, FusedClipAdam, GraphedTrainStep, fused_train_step

    world, rank, local = _dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the Multi-ATGCN path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world), file=sys.stderr)

    w = WORKLOADS[args.workload]
    per_gpu_batch = args.batch if args.batch else w["B"]
    micro = args.micro_batches
    if args.workload == "tract_8192":
        # BASELINE config 5: N = 8192, GLOBAL batch 512 over 8 GPUs = 64 samples per GPU, as four gradient-accumulation slices of 16
        # (a one-shot batch of 64 needs ~160 GB of saved activations in this layout); one set of workspaces only
        if not args.batch:
            per_gpu_batch = w["B"] // 8
        if micro <= 0:
            micro = 4
        args.no_graph_leg = True
    micro = max(micro, 1)
    cfg, df, _ = workload(args.workload, seed=0, batch=per_gpu_batch, device=dev)
    cfg["matgcn_mode"] = args.mode
    torch.manual_seed(0)
    model = MultiATGCN(dict(cfg), df).to(dev).train()
    broadcast_parameters(model)
    # executor:146-147 + 420-421: Adam(lr, eps) and clip_grad_norm_(5), fused over one flat bucket
    opt = FusedClipAdam(model.parameters(), lr=0.003, eps=1e-8, max_grad_norm=5.0)
    lib = _cabi.lib()

    def train_step(model, batch, opt, _bucket=None):
        return fused_train_step(model, batch, opt, micro_batches=micro)

    bucket = None

    # distinct host batches (pinned) so the e2e path really moves fresh data each step
    n_host = 4
    host = [make_batch(w["N"], per_gpu_batch, w["T_out"], seed=100 + rank * 16 + i, pin=True) for i in range(n_host)]
    resident = [{k: v.to(dev, non_blocking=True) for k, v in hb.items()} for hb in host]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------
    for i in range(args.warmup):
        train_step(model, resident[i % n_host], opt, bucket)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.matgcn_launch_count()
    tc0 = lib.matgcn_tc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss = train_step(model, resident[i % n_host], opt, bucket)
    e1.record()
    barrier()
    launches = lib.matgcn_launch_count() - l0
    tc_launches = lib.matgcn_tc_launch_count() - tc0
    ms_dev = e0.elapsed_time(e1) / args.steps
    last_loss = float(loss.item())
    # host time to ENQUEUE a step: three steps into an empty stream, clock stopped before any synchronisation (the launch queue
    # never fills, so this is pure host work: autograd, ctypes calls, tensor-map encodes, launches)
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    for i in range(3):
        train_step(model, resident[i % n_host], opt, bucket)
    host_ms = (time.perf_counter() - h0) * 1e3 / 3
    torch.cuda.synchronize()

    # ---- end-to-end timing: host buffers in, loss out, every step ------------------------------
    # Double-buffered: the pinned host batch of step i+1 is uploaded on a copy stream while step i computes, and the
    # loss of every step is copied to pinned host memory and read there (one step late, the last one before the clock
    # stops).  Every byte of every step's input crosses PCIe inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    loss_host = torch.zeros(args.steps + 8, dtype=torch.float32).pin_memory()

    def upload(i):
        with torch.cuda.stream(copy_stream):
            batch = {k: v.to(dev, non_blocking=True) for k, v in host[i % n_host].items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return batch, ev

    def e2e_run(n_steps, step_fn=None):
        if step_fn is None:
            step_fn = lambda b: train_step(model, b, opt, bucket)   # noqa: E731
        seen = []
        nxt = upload(0)
        prev_ev = None
        for i in range(n_steps):
            batch, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            if i + 1 < n_steps:
                nxt = upload(i + 1)
            loss = step_fn(batch)
            for v in batch.values():
                v.record_stream(torch.cuda.current_stream())
            loss_host[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            if prev_ev is not None:
                prev_ev.synchronize()
                seen.append(float(loss_host[i - 1]))
            prev_ev = done
        prev_ev.synchronize()
        seen.append(float(loss_host[n_steps - 1]))
        return seen

    e2e_run(min(max(args.warmup, 1), 3))
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_losses = e2e_run(args.steps)
    e3.record()
    barrier()
    assert len(e2e_losses) == args.steps and all(l == l for l in e2e_losses)
    ms_e2e = e2.elapsed_time(e3) / args.steps

    # ---- the same step with the batch gathered on the device from a series resident in HBM ----------
    # (the host sends B label-start indices per step instead of the assembled windows; reported next to e2e, not as it)
    tgen = torch.Generator().manual_seed(5 + rank)
    series = torch.randn(24 * 40, w["N"], 2, generator=tgen)
    series[..., 1] = ((torch.arange(24 * 40) % 24).float() / 24.0)[:, None]
    bank = DeviceWindowBank(series.to(dev), 24, w["T_out"], 2, 1, 1, 7, 28)
    valid = bank.valid_label_starts()
    if len(valid) >= per_gpu_batch:
        picks = [valid[torch.randperm(len(valid), generator=tgen)[:per_gpu_batch]].pin_memory() for _ in range(n_host)]
    else:
        picks = [valid[torch.randint(len(valid), (per_gpu_batch,), generator=tgen)].pin_memory() for _ in range(n_host)]
    for i in range(2):
        train_step(model, bank.assemble(picks[i % n_host]), opt, bucket)
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for i in range(args.steps):
        loss = train_step(model, bank.assemble(picks[i % n_host]), opt, bucket)
    float(loss.item())
    e5.record()
    barrier()
    ms_win = e4.elapsed_time(e5) / args.steps
    # ---- the same step captured once as a CUDA graph and replayed (train.GraphedTrainStep) -------------
    graph_leg = None
    if not args.no_graph_leg:
        gstep = GraphedTrainStep(model, opt, resident[0], input_slots=2)
        lg1 = lib.matgcn_launch_count()
        for i in range(max(args.warmup, 1)):
            gstep(resident[i % n_host])
        for s_ in range(len(gstep.graphs)):        # device-resident leg: the two input slots hold two resident batches
            gstep.load_batch(resident[s_ % n_host], slot=s_)
        barrier()
        e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e6.record()
        for i in range(args.steps):
            loss = gstep()
        e7.record()
        barrier()
        ms_graph = e6.elapsed_time(e7) / args.steps
        graph_loss = float(loss.item())
        eager_extra = lib.matgcn_launch_count() - lg1   # launches issued outside the graph while replaying (data parallel: the update)
        torch.cuda.synchronize()
        h0 = time.perf_counter()
        for i in range(3):
            gstep()
        host_ms_graph = (time.perf_counter() - h0) * 1e3 / 3
        torch.cuda.synchronize()
        def e2e_graph_run(n_steps):
            # the pinned host batch of step i+1 goes H2D on the copy stream STRAIGHT into the static input buffers of the graph that
            # will run step i+1 (two input slots, used alternately) while step i computes: no staging copy on the compute stream
            seen, done = [], [None] * len(gstep.graphs)

            def stage(i):
                slot = i % len(gstep.graphs)
                with torch.cuda.stream(copy_stream):
                    if done[slot] is not None:
                        copy_stream.wait_event(done[slot])      # the replay that last read these buffers has finished
                    gstep.load_batch(host[i % n_host], slot=slot)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                return slot, ev

            nxt = stage(0)
            prev_ev = None
            for i in range(n_steps):
                slot, ev = nxt
                torch.cuda.current_stream().wait_event(ev)
                if i + 1 < n_steps:
                    nxt = stage(i + 1)
                loss = gstep(slot=slot)
                loss_host[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
                d = torch.cuda.Event()
                d.record()
                done[slot] = d
                if prev_ev is not None:
                    prev_ev.synchronize()
                    seen.append(float(loss_host[i - 1]))
                prev_ev = d
            prev_ev.synchronize()
            seen.append(float(loss_host[n_steps - 1]))
            return seen

        e2e_graph_run(2)
        barrier()
        e8, e9 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e8.record()
        gl = e2e_graph_run(args.steps)
        e9.record()
        barrier()
        assert len(gl) == args.steps and all(l == l for l in gl)
        ms_graph_e2e = e8.elapsed_time(e9) / args.steps
        graph_leg = {"ms_per_step": ms_graph, "e2e_ms_per_step": ms_graph_e2e, "host_enqueue_ms_per_step": host_ms_graph,
                     "library_kernel_nodes": gstep.library_kernel_nodes, "loss": graph_loss,
                     "gpu_launches": int(gstep.library_kernel_nodes * args.steps + eager_extra * args.steps // (args.steps + max(args.warmup, 1)))}
        gstep.close()
        del gstep
    clocks = sampler.stop() if rank == 0 else {}
    strong = None
    if not args.no_strong_leg and args.workload == DEFAULT_WORKLOAD:
        del bank
        torch.cuda.empty_cache()
        strong = strong_scaling_leg(world, rank, dev, args.mode)

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e, ms_win] + ([graph_leg["ms_per_step"], graph_leg["e2e_ms_per_step"]] if graph_leg else []),
                         device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, ms_win = float(t[0]), float(t[1]), float(t[2])
        if graph_leg:
            graph_leg["ms_per_step"], graph_leg["e2e_ms_per_step"] = float(t[3]), float(t[4])
    global_batch = per_gpu_batch * world
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    # the roofline leg runs real train steps (with their all-reduce): EVERY rank takes part, rank 0 reports its own kernels' times
    peaks = _peaks()
    roof = None
    if model.matgcn_flags == 3 and per_gpu_batch <= 64:
        roof = recurrence_roofline(model, resident, opt, train_step, lib, w["N"], per_gpu_batch // micro, 5, 1, 24, peaks)
    # headline: the captured step when it was measured (it is the same work, issued as one cudaGraphLaunch), else the eager one
    eager = {"ms_per_step": ms_dev, "e2e_ms_per_step": ms_e2e, "host_enqueue_ms_per_step": host_ms, "gpu_launches": int(launches)}
    if graph_leg is not None and not args.eager_headline:
        ms_dev, ms_e2e, host_ms, launches, last_loss = (graph_leg["ms_per_step"], graph_leg["e2e_ms_per_step"],
                                                        graph_leg["host_enqueue_ms_per_step"], graph_leg["gpu_launches"], graph_leg["loss"])
    if rank == 0:
        if roof is None:   # modes / shapes that run one launch per phase: the support-propagation GEMM is the dominant kernel
            roof = propagation_roofline(w["N"], per_gpu_batch, 64, 4, model.ldm, dev, peaks, model.matgcn_flags)
        line = {"metric": METRIC, "value": global_batch / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": {0: "f32", 1: "tf32", 3: "bf16+tf32"}.get(model.matgcn_flags, "tf32"), "data": "synthetic",
                "config": {"workload": _workload_desc(args.workload, w, per_gpu_batch), "global_batch": global_batch,
                           "step": "zero_grad+forward+backward+allreduce+clip_grad_norm(5)+Adam (clip+Adam fused over one flat bucket)"
                                   + ("; captured once as a CUDA graph and replayed (train.GraphedTrainStep)"
                                      if graph_leg is not None and not args.eager_headline else ""),
                           "parallelism": "dp%d (batch-sharded, one flat-bucket all-reduce)" % world,
                           "micro_batches": micro,
                           "l2": "per-step working set (several GB of saved activations) is far larger than the 126 MB L2",
                           "mode": {0: "exact: fp32 FFMA kernels (1e-4 parity)",
                                    1: "fast: contractions on tcgen05 tensor cores as TF32, fp32 storage and accumulation",
                                    3: "fast: tcgen05 tensor cores; the streamed contractions (support propagation, per-node gate / "
                                       "candidate products and their reverse-step and weight-gradient counterparts) read bf16 "
                                       "operand twins, the rest is TF32; fp32 storage of the state, fp32 accumulation"}[model.matgcn_flags]},
                "clocks": clocks,
                "e2e": {"value": global_batch / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world,
                        "pipeline": "pinned host batch of step i+1 uploaded on a copy stream during step i (captured step: straight into "
                                    "the static input buffers of the next of two alternating graphs); every step's loss copied to pinned "
                                    "host memory and read there"},
                "e2e_device_windows": {"value": global_batch / (ms_win * 1e-3), "unit": UNIT, "ms_per_step": ms_win,
                                       "h2d_bytes_per_step": 8 * per_gpu_batch * world,
                                       "note": "SURVEY 8f f2: [T_total,N,F] series resident in HBM, matgcn_assemble_windows gathers "
                                               "each batch; the host uploads only the label-start indices"},
                "gpu_launches": int(launches), "tc_launches": int(tc_launches), "loss": last_loss, "roofline": roof,
                "host_enqueue_ms_per_step": host_ms}
        if args.workload == DEFAULT_WORKLOAD and per_gpu_batch == w["B"]:
            # SURVEY 8d, cfg 3: 2.77 TFLOP and 4.51 GB (bf16 storage) / 9.00 GB (fp32) x 3 per train step
            t_roof = max(2.77e12 / (peaks["bf16_tflops_sustained"] * 1e12), 3 * (4.51e9 if model.matgcn_flags == 3 else 9.00e9) / (peaks["hbm_gbs"] * 1e9))
            line["step_roofline"] = {"t_roof_ms": t_roof * 1e3, "frac": t_roof * 1e3 / ms_dev,
                                     "model": "SURVEY.md 8d: max(algorithmic FLOPs / sustained bf16 peak, algorithmic bytes / HBM peak) per step"}
        if graph_leg is not None:
            # one cudaGraphLaunch per step (step count, dropout key and learning rate live on the device)
            graph_leg["value"] = global_batch / (graph_leg["ms_per_step"] * 1e-3)
            graph_leg["e2e_value"] = global_batch / (graph_leg["e2e_ms_per_step"] * 1e-3)
            graph_leg["unit"] = UNIT
            eager["value"] = global_batch / (eager["ms_per_step"] * 1e-3)
            eager["e2e_value"] = global_batch / (eager["e2e_ms_per_step"] * 1e-3)
            eager["note"] = "the same step issued launch by launch (train.fused_train_step)"
            line["eager"] = eager
            graph_leg["note"] = ("train.GraphedTrainStep: zero_grad .. Adam captured once and replayed" +
                                 ("; data parallel: the graph ends after the backward, all-reduce + update follow eagerly" if world > 1 else ""))
            line["cuda_graph"] = graph_leg
        if strong is not None:
            line["strong_scaling"] = strong
        if world == 1 and model.matgcn_flags != 0 and not args.no_exact_leg:
            # the 1e-4-parity engine (fp32 FFMA kernels) on the same workload, same weights, a few steps: driver-run number
            # for the mode whose parity bound is the north star's fp32 one
            line["exact_mode"] = exact_mode_leg(cfg, df, model, resident, dev)
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = cpu_reference_run(args.workload, 2, 1, budget_s=args.cpu_budget if args.cpu_budget else 45.0)
                line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as exc:  # the baseline is informational; never lose the GPU line over it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: %r" % (exc,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def strong_scaling_leg(world, rank, dev, mode, steps=3, warmup=2):
    """BASELINE config 4 - the one scaling target the north star states: N = 883, 24 -> 12, GLOBAL batch 256 sharded over the
    ranks (256 / world samples per GPU), one flat-bucket all-reduce per step.  Returns (ms per step, max over ranks), per rank batch."""
    from multistgraph_b200.dp import broadcast_parameters
    from multistgraph_b200.model import MultiATGCN
    from multistgraph_b200.synthetic import WORKLOADS, make_batch, workload
    from multistgraph_b200.train import FusedClipAdam, fused_train_step

    w = WORKLOADS["pems07_scale"]
    per_rank = w["B"] // world
    cfg, df, _ = workload("pems07_scale", seed=0, batch=per_rank, device=dev)
    cfg["matgcn_mode"] = mode
    torch.manual_seed(0)
    model = MultiATGCN(dict(cfg), df).to(dev).train()
    broadcast_parameters(model)
    opt = FusedClipAdam(model.parameters(), lr=0.003, eps=1e-8, max_grad_norm=5.0)
    batches = [make_batch(w["N"], per_rank, w["T_out"], seed=300 + rank * 8 + i, device=dev) for i in range(2)]
    for i in range(warmup):
        fused_train_step(model, batches[i % 2], opt)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = fused_train_step(model, batches[i % 2], opt)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    last = float(loss.item())
    del opt, model, batches
    torch.cuda.empty_cache()
    return {"workload": "pems07_scale: N=883 nodes, GLOBAL batch 256 (%d per GPU), 24 steps in -> 12 out, K=5 supports" % per_rank,
            "value": w["B"] / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "n_gpus": world, "global_batch": w["B"],
            "per_gpu_batch": per_rank, "steps": steps, "warmup": warmup, "scaling": "strong", "loss": last,
            "roofline_samples_per_s": {"1_gpu_flop_bound": 8100.0, "8_gpu_bf16_weight_stream_bound_per_gpu": 7800.0,
                                       "source": "SURVEY.md section 8d work model"}}


def exact_mode_leg(cfg, df, fast_model, resident, dev, steps=3, warmup=1):
    """Train steps of the exact engine (matgcn_mode="exact": fp32 FFMA, parity bound 1e-4) on the bench workload."""
    from multistgraph_b200.model import MultiATGCN
    from multistgraph_b200.train import FusedClipAdam, fused_train_step

    c = dict(cfg)
    c["matgcn_mode"] = "exact"
    model = MultiATGCN(c, df).to(dev).train()
    model.load_state_dict(fast_model.state_dict())
    opt = FusedClipAdam(model.parameters(), lr=0.003, eps=1e-8, max_grad_norm=5.0)
    for i in range(warmup):
        fused_train_step(model, resident[i % len(resident)], opt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = fused_train_step(model, resident[i % len(resident)], opt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    batch = resident[0]["X"].shape[0]
    out = {"value": batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup, "dtype": "f32",
           "loss": float(loss.item()), "note": "same workload and weights through matgcn_mode='exact' (fp32 FFMA kernels; the mode "
                                               "held to max rel err 1e-4 against the oracle)"}
    del opt, model
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--mode", default="bf16", choices=["tf32", "bf16", "exact"],
                    help="bf16 (headline): tcgen05 tensor cores, bf16 operand twins for the streamed contractions, the rest "
                         "TF32; tf32: TF32 tensor cores with fp32 operands; exact: fp32 FFMA kernels (1e-4 parity)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--micro-batches", type=int, default=0,
                    help="gradient-accumulation slices per step (default 1; 4 for tract_8192, whose 64 samples per GPU do not fit one-shot)")
    ap.add_argument("--cpu-budget", type=float, default=0.0, help="seconds of CPU work allowed for the baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exact-leg", action="store_true", help="skip the few exact-mode steps reported as exact_mode")
    ap.add_argument("--no-graph-leg", action="store_true", help="skip the CUDA-graph replay of the step (cuda_graph key)")
    ap.add_argument("--eager-headline", action="store_true", help="report the launch-by-launch step as value / e2e")
    ap.add_argument("--no-strong-leg", action="store_true",
                    help="skip the strong-scaling leg (BASELINE config 4: N=883, global batch 256 sharded over the ranks)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
