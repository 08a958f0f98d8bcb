#!/usr/bin/env python
"""Benchmark of the multimodal-fusion block (forward + backward), BASELINE.json metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Workload (config.workload): BASELINE.json configs[1] - VisualBERT (visual_bert/pretrain) trunk, 12L/768/12h/3072,
128 text tokens + 100 regions x 2048 per sample, bf16 compute, dropout 0.1 (train mode), synthetic SampleList,
random-init weights.  A "step" is one forward+backward of the fusion block (region projection + embeddings +
12 layers; loss = fixed random projection of the sequence output) over one batch.

  value : samples/s with the step's inputs already resident in HBM (CUDA-event timed, max over ranks)
  e2e   : the same through the public SampleList API with HOST (pinned) inputs: H2D of ids/masks/features and a
          D2H read of the loss inside the timed region
  roofline   : dominant kernel (the tcgen05 GEMM at the FFN-up shape) timed alone with CUDA events,
               algorithmic FLOPs / duration vs MEASURED_PEAKS.json bf16 peak
  cpu_baseline: the oracle (CPU restatement of the reference, fp32, all host threads) on a bounded sample
  --impl reference : times that CPU arm alone, same metric / config / JSON shape.

N > 1 (torchrun): weak scaling, batch per GPU fixed, gradients of every rank averaged with NCCL all-reduce on
slices of the flat gradient buffer overlapped with the backward (mmf_b200.ddp).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_TOK, R_REG, FEAT, HID, HEADS, INTER, LAYERS, VOCAB = 128, 100, 2048, 768, 12, 3072, 12, 30522
S_LEN = T_TOK + R_REG
# algorithmic FLOPs per sample, fwd+bwd = 3 x fwd: 12 x (8 S H^2 + 4 S^2 H + 4 S H I) + 2 R F H
FWD_FLOPS = LAYERS * (8 * S_LEN * HID * HID + 4 * S_LEN * S_LEN * HID + 4 * S_LEN * HID * INTER) + 2 * R_REG * FEAT * HID
STEP_FLOPS_PER_SAMPLE = 3 * FWD_FLOPS


def model_config(p_drop):
    return types.SimpleNamespace(
        hidden_size=HID, num_attention_heads=HEADS, intermediate_size=INTER, num_hidden_layers=LAYERS,
        vocab_size=VOCAB, max_position_embeddings=512, type_vocab_size=2, visual_embedding_dim=FEAT,
        hidden_dropout_prob=p_drop, attention_probs_dropout_prob=p_drop, layer_norm_eps=1e-12, hidden_act="gelu",
        initializer_range=0.02)


def synthetic_sample_list(B, seed, pin):
    """SURVEY.md 8d synthetic inputs, on the HOST (pinned)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, VOCAB, (B, T_TOK), generator=g)
    lens = torch.randint(T_TOK // 2, T_TOK + 1, (B,), generator=g)
    mask = (torch.arange(T_TOK)[None, :] < lens[:, None]).long()
    seg = torch.zeros(B, T_TOK, dtype=torch.long)
    feats = torch.randn(B, R_REG, FEAT, generator=g).abs()
    maxf = torch.randint(R_REG // 2, R_REG + 1, (B,), generator=g)
    sl = {"input_ids": ids, "input_mask": mask, "segment_ids": seg, "image_feature_0": feats,
          "image_info_0": {"max_features": maxf}}
    if pin:
        sl = {k: (v.pin_memory() if hasattr(v, "pin_memory") else {kk: vv.pin_memory() for kk, vv in v.items()})
              for k, v in sl.items()}
    return sl


def to_device(sl, dev):
    out = {}
    for k, v in sl.items():
        out[k] = {kk: vv.to(dev, non_blocking=True) for kk, vv in v.items()} if isinstance(v, dict) else v.to(
            dev, non_blocking=True)
    return out


def h2d_bytes(sl):
    n = 0
    for v in sl.values():
        for t in (v.values() if isinstance(v, dict) else [v]):
            n += t.numel() * t.element_size()
    return n


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)"""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 1590.0, 1400.0, "fallback"


# --------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's path, oracle/fusion_oracle.py) on the host cores
# --------------------------------------------------------------------------------------------------------
def cpu_step_fn(B, p_drop):
    import torch
    from oracle import fusion_oracle as O
    torch.manual_seed(0)
    sd = O.make_encoder_weights(LAYERS, HID, INTER, seed=0, prefix="encoder.layer")
    g = torch.Generator().manual_seed(1)
    sd["emb.word_embeddings.weight"] = torch.randn(VOCAB, HID, generator=g) * 0.02
    sd["emb.position_embeddings.weight"] = torch.randn(512, HID, generator=g) * 0.02
    sd["emb.token_type_embeddings.weight"] = torch.randn(2, HID, generator=g) * 0.02
    sd["emb.token_type_embeddings_visual.weight"] = torch.randn(2, HID, generator=g) * 0.02
    sd["emb.position_embeddings_visual.weight"] = torch.randn(512, HID, generator=g) * 0.02
    sd["emb.projection.weight"] = torch.randn(HID, FEAT, generator=g) * 0.02
    sd["emb.projection.bias"] = torch.zeros(HID)
    sd["emb.LayerNorm.weight"] = torch.ones(HID)
    sd["emb.LayerNorm.bias"] = torch.zeros(HID)
    for v in sd.values():
        v.requires_grad_(True)
    sl = synthetic_sample_list(B, 1234, pin=False)
    w_rand = torch.randn(B, S_LEN, HID, generator=g)

    def step():
        for v in sd.values():
            v.grad = None
        image_mask, vtype, att = O.visual_bert_masks(sl["input_mask"], sl["image_info_0"]["max_features"], R_REG)
        keep = None
        masks = None
        if p_drop > 0:   # nn.Dropout-style RNG dropout at the reference's four sites per layer
            keep = "rng"
            masks = [{"attn": "rng", "self_out": "rng", "out": "rng"} for _ in range(LAYERS)]
        emb = O.visio_linguistic_embeddings(sl["input_ids"], sl["segment_ids"], sl["image_feature_0"], vtype, sd, "emb",
                                            keep=keep, p=p_drop)
        out = O.bert_encoder(emb, O.extended_attention_mask(att), sd, "encoder", LAYERS, HEADS, masks, p_drop, p_drop)
        loss = (out * w_rand).sum()
        loss.backward()
        return float(loss.detach())
    return step


def usable_cores():
    """host threads this process may actually use: affinity mask capped by the cgroup CPU quota"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:
            quota, period = fh.read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def run_cpu_arm(steps, warmup, B, p_drop):
    import torch
    cores = usable_cores()
    torch.set_num_threads(cores)
    step = cpu_step_fn(B, p_drop)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return B / med, med * 1e3, torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("MMFB_BENCH_BATCH", "166")), help="samples per GPU")
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="device-resident steps only (for ncu launch lists)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "VisualBERT visual_bert/pretrain trunk 12L/768/12h/3072, 128 tokens + 100 regions x 2048 "
                          "(BASELINE.json configs[1]); fwd+bwd of region projection + embeddings + 12 fusion layers",
              "batch_per_gpu": args.batch, "global_batch": args.batch * max(world, 1), "seq_len": S_LEN,
              "dropout": args.dropout, "parallelism": "dp%d" % max(world, 1),
              "batch_choice": "166 samples x 228 tokens = 147.8 -> 148 pair tiles of 256 rows: every GEMM of the block is "
                              "an exact number of waves on 148 SMs (sweep in profiles/README.md)",
              "l2_policy": "activations + weights touched per step (~%d MB at batch %d) exceed the 126 MB L2" % (
                  int(args.batch * 72), args.batch)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cpu_steps = max(1, min(args.steps, 5))
        v, ms, cores = run_cpu_arm(cpu_steps, max(1, min(args.warmup, 1)), args.cpu_batch, args.dropout)
        line = {"impl": "reference", "metric": "multimodal-fusion samples/sec (fwd+bwd)", "value": v, "unit": "samples/s",
                "n_gpus": args.gpus, "steps": cpu_steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                                 "sample": "batch %d of the same workload per step, fp32, oracle/fusion_oracle.py" % args.cpu_batch},
                "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from mmf_b200 import functional as F, lib
    from mmf_b200.visual_bert import B200VisualBERT

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = B200VisualBERT(model_config(args.dropout)).to(dev).train()
    ddp = None
    if world > 1:
        from mmf_b200.ddp import B200DataParallel
        ddp = B200DataParallel(model)
    B = args.batch
    host_sl = synthetic_sample_list(B, 1234 + rank, pin=True)
    dev_sl = to_device(host_sl, dev)
    w_rand = torch.randn(B, S_LEN, HID, device=dev)

    def step(sl):
        model.zero_grad(set_to_none=True)
        out = (ddp(sl) if ddp is not None else model(sl))["sequence_output"]
        loss = (out * w_rand).sum()
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(dev_sl)
    barrier()
    # ---------------- device-resident timing ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profile:
        torch.cuda.profiler.start()    # ncu --profile-from-start off: capture exactly the timed steps
    e0.record()
    for _ in range(args.steps):
        step(dev_sl)
    e1.record()
    barrier()
    if args.profile:
        torch.cuda.profiler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_total / args.steps, "launches": launches}))
        return 0
    # ---------------- end-to-end timing (host inputs, loss read back) ----------------
    for _ in range(2):
        float(step(to_device(host_sl, dev)).detach())
    barrier()
    # Input pipeline as in the reference trainer (pinned SampleList + non_blocking copies, sample.py:326-370): the H2D of
    # step i+1 is issued on a copy stream while step i computes.  Every step still copies its full inputs H2D and reads
    # its loss back D2H inside the timed region.
    copy_stream = torch.cuda.Stream()

    def prefetch():
        with torch.cuda.stream(copy_stream):
            batch = to_device(host_sl, dev)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return batch, ev

    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    nxt = prefetch()
    for i in range(args.steps):
        batch, ev = nxt
        torch.cuda.current_stream().wait_event(ev)
        if i + 1 < args.steps:
            nxt = prefetch()
        loss = step(batch)
        for v in batch.values():      # keep the copy-stream allocations alive until the compute stream has used them
            for t in (v.values() if isinstance(v, dict) else [v]):
                t.record_stream(torch.cuda.current_stream())
        _ = float(loss.detach())   # D2H read of the step's result
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)
    e2e_value = B * world / (ms_e2e / args.steps / 1e3)

    line = None
    if rank == 0:
        burst, sustained, how = peaks()
        # dominant kernel alone: the tcgen05 GEMM at the FFN-up shape [B*S, 768] x [3072, 768]^T (+bias+GELU epilogue)
        M = B * S_LEN
        a = torch.randn(M, HID, device=dev).to(torch.bfloat16)
        w = (torch.randn(INTER, HID, device=dev) * 0.02).to(torch.bfloat16)
        bias = torch.zeros(INTER, device=dev, dtype=torch.bfloat16)
        o1 = torch.empty(M, INTER, device=dev, dtype=torch.bfloat16)
        o2 = torch.empty_like(o1)
        flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
        times = []
        for i in range(13):
            flush.zero_()   # L2 flush between timed launches
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            F.gemm(a, w, epi=lib.EPI_BIAS_GELU, bias=bias, out=o1, out2=o2)
            g1.record()
            torch.cuda.synchronize()
            if i >= 3:
                times.append(g0.elapsed_time(g1))
        k_ms = statistics.mean(times)
        k_flops = 2.0 * M * HID * INTER
        achieved = k_flops / (k_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "gemm_kernel<256,K-major,K-major,BIAS_GELU> FFN-up [%d,768]x[3072,768]^T" % M,
                    "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                    "peak_source": "%s bf16_tflops (burst: kernel timed alone)" % how,
                    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at batch 166, one `ncu --set full` capture
                    # (profiles/r1_ncu_dominant_b166.txt): 63.1 MB + 405.5 MB vs 62.8 + 465.1 MB algorithmic
                    "traffic": 468.6e6 if B == 166 else None,
                    "kernel_ms": k_ms, "flops_per_launch": k_flops,
                    "step_tflops": STEP_FLOPS_PER_SAMPLE * value / 1e12 / world,
                    "step_frac_of_sustained": STEP_FLOPS_PER_SAMPLE * value / 1e12 / world / sustained}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            v, ms, cores = run_cpu_arm(3, 1, args.cpu_batch, args.dropout)
            cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": "3 timed steps of batch %d of the same workload, fp32, oracle/fusion_oracle.py" % args.cpu_batch}
        line = {"metric": "multimodal-fusion samples/sec (fwd+bwd)", "value": value, "unit": "samples/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config,
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes(host_sl),
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
