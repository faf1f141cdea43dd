#!/usr/bin/env python
"""bench.py -- site-saturation ddG throughput of the CUDA path (and of the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl ours|reference]

One "step" = one pass of the whole hot path (k-NN graph -> edge features -> 3 Enc + 3 Dec layers -> head table
for all 21 letters) over one batch of synthetic proteins, through the product's own entry points.

  N = 1   workload c2 = BASELINE.json configs[1]: one 300-residue synthetic protein, 6000 residue-mutations per step.
  N > 1   the SAME unit per GPU (weak scaling): a batch of N 300-residue proteins, LPT-sharded over the ranks by
          thermompnn_b200.sharding (one protein each), every rank's tables written into its slot of one flat device
          buffer and exchanged with ONE NCCL all_gather_into_tensor inside the timed step (sharding.ShardedScan, the
          product's multi-GPU path).  The line also carries "sharded_c5": BASELINE configs[4] (64 x 2000 residues)
          LPT-sharded over the same N ranks through the same path (strong scaling: total work fixed).
  c3 / c4 / c5 are the other BASELINE configs (256 proteins of 150..500 / the Megascale-shaped 298 proteins / the
          2000-residue stress proteins), selected with --workload.

Keys (one JSON line on rank 0's stdout):
  value      mutations/s, inputs already resident in HBM (device pointers into the C ABI), CUDA events on the launching
             stream, max over ranks, L2 flushed (256 MB write) between steps
  e2e        the same metric through the C ABI with pinned HOST buffers: H2D of the packed inputs, the kernel chain, (N > 1:
             the all-gather,) D2H of the result tables -- all inside the timed region (wall clock per step, max over ranks)
  roofline   the kernel with the largest share of the step: CUDA-event duration per launch (engine profiling hooks)
             against the measured peaks of MEASURED_PEAKS.json; "large_batch" repeats this for workload c3, where the
             kernels are throughput- rather than latency-bound
  cpu_baseline / --impl reference: the oracle's restatement of the reference's CPU path timed on this box's host cores,
             (i) as the reference runs it (per-mutation head loop, all 20 L iterations executed) and (ii) the same
             ProteinMPNN pass with a vectorised head (the generous baseline)
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

# stdout carries exactly ONE JSON line.  Libraries write to the C-level stdout too (NCCL prints its version banner with
# printf when NCCL_DEBUG is set on the box), so file descriptor 1 is pointed at stderr for the whole run and the
# result line is written to the saved descriptor at the end.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    os.write(_RESULT_FD, (json.dumps(obj) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CKPT = os.path.join(ROOT, "tests", "golden", "thermompnn_default_slim.npz")
MEGA = os.path.join(ROOT, "tests", "golden", "megascale_shape.npz")
METRIC = "residue-mutations/sec (site-saturation ddG)"
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # nominal: 148 SMs x 128 FP32 FMA lanes x 2 flop x 1.965 GHz = 74.4

# algorithmic work of the reference formulation, per kernel class, per edge / per residue
ALGO = {  # name: (flop_per_edge, flop_per_node, bytes_per_edge, bytes_per_node)
    "edge_feat": (2 * 416 * 128 + 2 * 128 * 128, 0, 512 + 4, 15 * 4),
    "enc_msg": (2 * (384 + 128 + 128) * 128, 0, 512 + 4, 1024),
    "dec_msg": (2 * (512 + 128 + 128) * 128, 0, 512 + 4, 1024),
    "edge_upd": (2 * (384 + 128 + 128) * 128, 0, 1024 + 4, 512),
    "node_update": (0, 2 * 2 * 128 * 512, 0, 1024 + 4 * 512),
    "head": (0, 349504, 0, 3 * 512 + 84),
    "knn": (0, 0, 8, 64 + 112),
}
# FLOPs the kernels actually execute (h_V projections hoisted per residue, W3 after the neighbour sum):
EXEC = {"edge_feat": 2 * 416 * 128 + 2 * 128 * 128, "enc_msg": 2 * 2 * 128 * 128, "dec_msg": 2 * 2 * 128 * 128,
        "edge_upd": 3 * 2 * 128 * 128, "node_update": 2 * (128 * 128 + 2 * 128 * 512 + 4 * 128 * 128), "head": 349504}


def base_name(k):
    return k[3:] if k.startswith("tc_") else k


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def make_workload(name, world=1):
    """(proteins, description, listed mutations or None) of the WHOLE job (all ranks)."""
    from thermompnn_b200 import synth
    if name == "c2":
        if world == 1:
            return [synth.synth(300, seed=0)], "BASELINE configs[1]: one 300-residue synthetic protein, full 20-AA scan", None
        return ([synth.synth(300, seed=r) for r in range(world)],
                f"BASELINE configs[1] x {world}: a batch of {world} 300-residue synthetic proteins (one per GPU, LPT-sharded), full 20-AA scan", None)
    if name == "c3":
        return synth.config3(), "BASELINE configs[2]: 256 synthetic proteins, lengths U[150,500] (seed 1234)", None
    if name == "c4":
        pdbs, prot, pos, mut = synth.config4(MEGA)
        return pdbs, "BASELINE configs[3]: Megascale-shaped batch, 298 proteins L 32..72, synthetic coordinates seeded by crc32(WT_name)", len(prot)
    if name == "c5":
        return synth.config5(), "BASELINE configs[4]: 64 synthetic proteins of 2000 residues", None
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed regions run.  ONE sampler (rank 0) polls every
    GPU of the job: eight processes polling NVML every few milliseconds contend on driver locks with the kernel launches."""

    def __init__(self, indices, period=0.02):
        super().__init__(daemon=True)
        self.indices, self.period = list(indices), period
        self.samples, self.marks, self.reasons, self.stop_flag, self.max_mhz = [], [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in self.indices]
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h[0], pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            for h in self.h:
                try:
                    self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
                except Exception:
                    pass
            time.sleep(self.period)

    def result(self, region):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        allv = [v for _, v in self.samples]
        inr = [v for t, v in self.samples if region[0] <= t <= region[1]]
        return {"sm_mhz": statistics.median(inr or allv) if allv else None, "sm_max_mhz": self.max_mhz,
                "samples_in_value_region": len(inr), "samples_under_load": len(allv), "sm_mhz_all_legs": statistics.median(allv) if allv else None,
                "legs_sampled": f"value (device-timed), e2e and per-kernel profile legs of the main workload, every {1e3 * self.period:.0f} ms",
                "gpus_sampled": len(self.indices), "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's port of TransferModel.forward on host cores
# ---------------------------------------------------------------------------------------------
def reference_weights():
    import torch
    from oracle import thermompnn_oracle as O
    w = O.cast_weights(dict(np.load(CKPT)))
    # The per-mutation loop needs the full k=9 convolutions.  Off-centre taps and the attention branch cannot
    # influence the result on a length-1 sequence (tests/test_oracle.py), so they are filled with noise here.
    g = torch.Generator().manual_seed(0)
    full = torch.randn(384, 384, 9, generator=g) * 0.02
    full[:, :, 4] = w["light_attention.feature_convolution.weight_center"]
    w["light_attention.feature_convolution.weight"] = full
    w["light_attention.attention_convolution.weight"] = torch.randn(384, 384, 9, generator=g) * 0.02
    w["light_attention.attention_convolution.bias"] = torch.zeros(384)
    return w


def reference_pass(w, pdb, n_head=None):
    """The reference CPU path on one protein, as upstream runs it: the ProteinMPNN pass, then one head evaluation per
    mutation (transfer_model.py:86-120).  n_head = None executes ALL 20 L iterations (a measurement); an integer runs that
    many (thread calibration only).  Returns (seconds, n_mutations, mpnn seconds)."""
    import torch
    from oracle import thermompnn_oracle as O
    X, S, mask, ridx, cenc = O.featurize(pdb)
    L = len(S)
    muts = [(i, a) for i in range(L) if pdb["seq"][i] != "-" for a in range(20)]
    if n_head is not None:
        muts = muts[:: max(1, len(muts) // n_head)][:n_head]
    with torch.no_grad():
        t0 = time.perf_counter()
        hidden, h_S = O.mpnn_forward(w, torch.from_numpy(X), torch.from_numpy(S), torch.from_numpy(mask),
                                     torch.from_numpy(ridx), torch.from_numpy(cenc))
        t1 = time.perf_counter()
        for p, a in muts:
            O.head_per_mutation(w, hidden, h_S, p, int(S[p]), a)
        t2 = time.perf_counter()
    return t2 - t0, len(muts), t1 - t0


def vectorised_pass(w, pdb):
    """The generous CPU baseline: the same ProteinMPNN pass + ONE [L,384] head evaluation for all mutations."""
    import torch
    from oracle import thermompnn_oracle as O
    t0 = time.perf_counter()
    with torch.no_grad():
        O.ssm_table(w, pdb)
    return time.perf_counter() - t0, 20 * sum(a != "-" for a in pdb["seq"])


def choose_threads(w, pdb):
    """The reference leaves threading to torch's default (= all cores), which collapses on many-core hosts because its
    ops are tiny (conv1d over a length-1 sequence, 128-wide GEMMs).  Be fair to the CPU: try a few thread counts on a
    short sample and keep the fastest."""
    import torch
    ncpu = os.cpu_count() or 1
    best, best_t = 1, None
    for nt in [c for c in (1, 2, 4, 8, 16, 32, 64) if c <= ncpu]:
        torch.set_num_threads(nt)
        reference_pass(w, pdb, 20)                       # warm
        t = reference_pass(w, pdb, 60)[0]
        if best_t is None or t < best_t:
            best, best_t = nt, t
    torch.set_num_threads(best)
    return best


def cpu_baseline(pdb, repeats=3):
    import torch
    w = reference_weights()
    cores = choose_threads(w, pdb)
    runs = [reference_pass(w, pdb) for _ in range(repeats)]
    t_as_is = statistics.median(r[0] for r in runs)
    n_mut = runs[0][1]
    vruns = [vectorised_pass(w, pdb) for _ in range(repeats + 1)][1:]
    t_vec = statistics.median(r[0] for r in vruns)
    return {"value": n_mut / t_as_is, "unit": "mutations/s", "cores": cores, "kind": "port",
            "sample": f"{repeats} full passes over one protein of the workload ({n_mut} mutations): ProteinMPNN pass + ALL {n_mut} per-mutation head "
                      f"iterations executed (nothing extrapolated); median; thread count = fastest of {{1,2,4,...,64}} <= {os.cpu_count()} logical CPUs",
            "as_is": {"value": n_mut / t_as_is, "s_per_protein": t_as_is, "mpnn_s": statistics.median(r[2] for r in runs)},
            "vectorised_head": {"value": n_mut / t_vec, "s_per_protein": t_vec,
                                "note": "same ProteinMPNN pass, head evaluated once for all positions ([L,384] matmul chain): the generous CPU baseline"},
            "torch_threads": torch.get_num_threads()}, w


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    pdbs, desc, _ = make_workload(args.workload, world)
    pdb = pdbs[0]
    w = reference_weights()
    cores = choose_threads(w, pdb)
    times, n_mut = [], 0
    for it in range(args.warmup + args.steps):
        t, n_mut, _ = reference_pass(w, pdb)
        if it >= args.warmup:
            times.append(t)
    ms = 1e3 * sum(times) / len(times)
    value = n_mut / (ms / 1e3)
    tv = statistics.median(vectorised_pass(w, pdb)[0] for _ in range(3))
    sample = (f"per step: ONE protein of the workload, the reference CPU path in full -- ProteinMPNN pass + all {n_mut} per-mutation head "
              f"iterations executed (no extrapolation); the CPU's mutations/s does not depend on how many proteins a step holds")
    emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "mutations/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "host": "cpu", "proteins_per_step": 1},
        "cpu_baseline": {"value": value, "unit": "mutations/s", "cores": cores, "kind": "port", "sample": sample,
                         "vectorised_head": {"value": n_mut / tv, "s_per_protein": tv}},
        "e2e": {"value": value, "unit": "mutations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
class Job:
    """One sharded workload on this process group: device-timed and end-to-end-timed steps through sharding.ShardedScan."""

    def __init__(self, model, pdbs, dev, stream):
        import torch
        from thermompnn_b200.sharding import ShardedScan
        self.torch = torch
        self.scan = ShardedScan(model, pdbs, dev)
        self.stream = stream
        self.world, self.rank = self.scan.world, self.scan.rank
        self.h_in = self.scan.host_inputs() if self.scan.packed is not None else None
        self.d_in = self.scan.device_inputs() if self.scan.packed is not None else None
        p = self.scan.packed
        self.N = p.n_residues if p is not None else 0
        self.B = p.n_proteins if p is not None else 0
        self.n_edges = int(sum(min(48, int(l)) * int(l) for l in np.diff(p.offsets))) if p is not None else 0
        self.total_residues = sum(self.scan.plan.lengths)

    def step_device(self):
        self.scan.compute_local(self.d_in, stream=self.stream.cuda_stream, device_inputs=True)
        self.scan.gather()

    def step_host(self):
        self.scan.compute_local(self.h_in, stream=self.stream.cuda_stream)
        self.scan.gather()
        self.scan.host.copy_(self.scan.flat, non_blocking=True)
        self.stream.synchronize()

    def bytes_per_step(self):
        h2d = 4 * (self.B + 1) + self.N * (48 + 4 + 4 + 4 + 1)
        d2h = self.world * self.scan.plan.max_rows * 21 * 4
        return h2d, d2h


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from thermompnn_b200.transfer_model import get_trained_model, TransferModel
    from thermompnn_b200.weights import random_state

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    real_weights = os.path.exists(CKPT)
    model = (get_trained_model(CKPT, override_custom=True) if real_weights else TransferModel(state=random_state(0))).eval().to(dev)
    eng = model._get_engine()

    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    st = torch.cuda.Stream(device=dev)
    W = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(job, steps, profile=True):
        """(ms_per_step device-timed, e2e ms per step, launches per step, per-kernel profile, wall region) for one job."""
        with torch.cuda.stream(st):
            for _ in range(W):
                flush_buf.zero_()
                job.step_device()
            barrier()
            for _ in range(W):
                job.step_host()
            barrier()
            # ---- value: device-resident inputs, CUDA events per step, L2 flushed between steps
            eng.profile_reset()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            barrier()
            t_w0 = time.perf_counter()
            for a, b in ev:
                flush_buf.zero_()
                a.record(st)
                job.step_device()
                b.record(st)
            barrier()
            t_w1 = time.perf_counter()
            launches = eng.launch_count()
            graph_steps = eng.graph_launch_count()
            dev_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)) / steps
            # ---- e2e: pinned host buffers through the C ABI, copies and the collective inside the timed region
            e2e_t = []
            barrier()
            for _ in range(steps):
                flush_buf.zero_()
                barrier()
                t0 = time.perf_counter()
                job.step_host()
                e2e_t.append(time.perf_counter() - t0)
            barrier()
            e2e_ms = max_over_ranks(1e3 * sum(e2e_t) / len(e2e_t))
            prof = None
            if profile:
                # ---- per-kernel CUDA-event durations (direct launches, events around every kernel), L2 flushed
                eng.set_profiling(True)
                eng.profile_reset()
                n = min(steps, 20)
                for _ in range(n):
                    flush_buf.zero_()
                    job.scan.compute_local(job.d_in, stream=st.cuda_stream, device_inputs=True)
                torch.cuda.synchronize(dev)
                prof = (eng.profile(), n)
                eng.set_profiling(False)
        return dev_ms, e2e_ms, launches // steps, graph_steps, prof, (t_w0, t_w1)

    def kernel_table(job, prof, n):
        tot = sum(v[0] for v in prof.values())
        kern = {}
        for k, (ms, cnt) in prof.items():
            per = ms / max(cnt, 1)
            fe, fn, be, bn = ALGO[base_name(k)]
            kern[k] = {"us_per_launch": 1e3 * per, "launches_per_step": cnt // n, "share": ms / tot,
                       "algo_tflops": (fe * job.n_edges + fn * job.N) / (per * 1e-3) / 1e12,
                       "algo_gbs": (be * job.n_edges + bn * job.N) / (per * 1e-3) / 1e9}
        return kern

    def roofline(job, kern, prof, peaks, traffic_tab):
        def roof(k):
            tk = kern[k]
            alg = base_name(k)
            flops = ALGO[alg][0] * job.n_edges + ALGO[alg][1] * job.N
            r = {"kernel": k, "us_per_launch": tk["us_per_launch"], "share_of_step": tk["share"], "traffic": traffic_tab.get(k),
                 "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture "
                                 "(profiles/r02_traffic.json); null when no capture exists for this workload"}
            hbm = {"achieved": tk["algo_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": tk["algo_gbs"] / peaks["hbm_gbs"],
                   "peak_source": peaks["source"] + ", HBM copy burst"}
            if not flops:
                r.update({"bound": "hbm", **hbm, "note": "FP32 / integer SIMT kernel"})
                return r
            if k.startswith("tc_"):            # tcgen05 kernels
                ex = EXEC[alg] * (job.n_edges if ALGO[alg][0] else job.N) / (tk["us_per_launch"] * 1e-6) / 1e12
                r.update({"bound": "tensor", "achieved": tk["algo_tflops"], "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                          "frac": tk["algo_tflops"] / peaks["bf16_tflops"], "peak_source": peaks["source"] + ", dense bf16 burst",
                          "tensor_pipe_model": {"executed_fp32_equiv_tflops": ex, "mma_tflops_issued": 3 * ex, "tf32_peak": peaks["bf16_tflops"] / 2,
                                                "frac_of_tf32_peak": 3 * ex / (peaks["bf16_tflops"] / 2)},
                          "hbm": hbm,
                          "note": "tcgen05 kind::tf32 with the 3xTF32 split (fp32-accurate): 3 tensor-core MACs per executed MAC at half the bf16 "
                                  "rate, so frac <= 1/6 of the dense bf16 peak by construction; `achieved` counts the reference formulation's "
                                  "FLOPs (SURVEY 8d), tensor_pipe_model the MMAs actually issued against the tf32 peak (= bf16 / 2)"})
            else:
                r.update({"bound": "fp32", "achieved": tk["algo_tflops"], "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                          "frac": tk["algo_tflops"] / FP32_PEAK_TFLOPS,
                          "peak_source": "nominal FP32 FMA peak: 148 SMs x 128 lanes x 2 flop x 1.965 GHz (no measured FP32 figure in MEASURED_PEAKS.json)",
                          "frac_of_bf16_tensor_peak": tk["algo_tflops"] / peaks["bf16_tflops"], "hbm": hbm,
                          "note": "per-residue kernel on the FP32 / warp-level (mma.sync) pipes: a latency chain at this batch size"})
            return r
        top = max(kern, key=lambda k: prof[k][0])
        out = roof(top)
        edge = [k for k in kern if base_name(k) in ("edge_upd", "enc_msg", "dec_msg", "edge_feat")]
        top_edge = max(edge, key=lambda k: prof[k][0])
        out["top_edge_kernel"] = roof(top_edge) if top_edge != top else None
        return out

    # ------------------------------------------------------------------------------------------------
    pdbs, desc, n_listed = make_workload(args.workload, world)
    job = Job(model, pdbs, dev, st)
    n_mut_total = 20 * job.total_residues                 # synthetic proteins have no gaps: 20 substitutions per residue
    # one GPU: every 5 ms (a 100-step timed region is ~30 ms); several GPUs: ONE sampler polls all of them every 20 ms
    period = float(os.environ.get("TMPNN_BENCH_SAMPLE_PERIOD", "0.005" if world == 1 else "0.02"))
    sampler = ClockSampler(range(world), period) if (rank == 0 and period > 0) else None
    if sampler:
        sampler.start()
    dev_ms, e2e_ms, launches_per_step, graph_steps, (prof, prof_n), region = measure(job, args.steps)
    clocks = sampler.result(region) if sampler else None      # (stopped here: the companion legs below have idle CPU phases)
    value = n_mut_total / (dev_ms / 1e3)
    e2e_value = n_mut_total / (e2e_ms / 1e3)
    h2d, d2h = job.bytes_per_step()
    peaks = load_peaks()

    extra = {}
    # ---- strong-scaling companion for N > 1: BASELINE configs[4] sharded over the same ranks through the same path
    if world > 1 and args.workload == "c2" and not args.no_extra:
        from thermompnn_b200.sharding import ShardPlan
        from thermompnn_b200 import synth
        plan = ShardPlan([2000] * 64, world, rank)
        mine = set(plan.mine)
        # only this rank's proteins are generated; the others are placeholders of the right LENGTH for the plan
        c5 = [synth.synth(2000, seed=2000 + i) if i in mine else {"seq": "A" * 2000} for i in range(64)]
        job5 = Job(model, c5, dev, st)
        s5 = max(3, min(args.steps, 10))
        d5, e5, l5, _, _, _ = measure(job5, s5, profile=False)
        n5 = 20 * 64 * 2000
        extra["sharded_c5"] = {"workload": "BASELINE configs[4]: 64 synthetic proteins of 2000 residues, LPT-sharded over the ranks, one all_gather_into_tensor "
                                           "of the [L,21] tables per step", "scaling": "strong", "proteins_per_gpu": job5.B, "residues_per_gpu": job5.N,
                               "steps": s5, "ms_per_step": d5, "value": n5 / (d5 / 1e3), "unit": "mutations/s",
                               "e2e": {"ms_per_step": e5, "value": n5 / (e5 / 1e3)}, "launches_per_step": l5}
        del job5
    # ---- large-batch roofline at N = 1: workload c3 (the kernels' throughput regime), a few steps
    if world == 1 and args.workload == "c2" and not args.no_extra:
        from thermompnn_b200 import synth
        job3 = Job(model, synth.config3(), dev, st)
        s3 = 5
        d3, e3, l3, _, (prof3, n3), _ = measure(job3, s3)
        kern3 = kernel_table(job3, prof3, n3)
        tt = {}
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tp):
            tt = json.load(open(tp)).get("c3", {})
        extra["large_batch"] = {"workload": "BASELINE configs[2]: 256 synthetic proteins, lengths U[150,500] (seed 1234), one call", "steps": s3,
                                "residues": job3.N, "edges": job3.n_edges, "ms_per_step": d3, "value": 20 * job3.N / (d3 / 1e3),
                                "e2e": {"ms_per_step": e3, "value": 20 * job3.N / (e3 / 1e3)}, "unit": "mutations/s",
                                "roofline": roofline(job3, kern3, prof3, peaks, tt),
                                "kernels": {k: {kk: v[kk] for kk in ("us_per_launch", "launches_per_step", "share", "algo_tflops", "algo_gbs")} for k, v in kern3.items()}}
        del job3

    if rank == 0:
        kern = kernel_table(job, prof, prof_n)
        traffic_tab = {}
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tp):
            traffic_tab = json.load(open(tp)).get(args.workload, {})
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu, _ = cpu_baseline(pdbs[0])
        out = {
            "metric": METRIC, "value": value, "unit": "mutations/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic coordinates (thermompnn_b200.synth), " + ("reference checkpoint weights" if real_weights else "random-init weights"),
            "config": {"workload": desc, "proteins_total": len(pdbs), "proteins_per_gpu": job.B, "residues_per_gpu": job.N,
                       "edges_per_gpu": job.n_edges, "k_neighbors": 48, "hidden": 128, "l2": "flushed between timed steps (256 MB write)",
                       "parallelism": f"dp{world}",
                       "multi_gpu": ("LPT shard by L*min(48,L) (thermompnn_b200.sharding.ShardedScan), engine writes into its slot of one flat device "
                                     "buffer, ONE in-place all_gather_into_tensor (NCCL) per step inside the timed region") if world > 1 else None},
            "e2e": {"value": e2e_value, "unit": "mutations/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "includes": "H2D of packed inputs, kernel chain" + (", all_gather_into_tensor" if world > 1 else "") + ", D2H of the tables"},
            "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
            "graph_replayed_steps": int(graph_steps),
            "roofline": roofline(job, kern, prof, peaks, traffic_tab), "cpu_baseline": cpu, "clocks": clocks, "kernels": kern,
            "wall_s_timed_region": region[1] - region[0],
        }
        if n_listed:
            out["listed_mutations"] = {"count": n_listed, "value": n_listed / (dev_ms / 1e3), "e2e_value": n_listed / (e2e_ms / 1e3), "unit": "mutations/s",
                                       "note": "the 271 231 single substitutions the Megascale CSVs list (datasets.py:66-73 filter); `value` counts the full scan"}
        out.update(extra)
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the companion measurements (large_batch at N=1, sharded_c5 at N>1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
