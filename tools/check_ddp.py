"""torchrun --nproc-per-node 2 tools/check_ddp.py : gradients after B200DataParallel == mean of the two ranks' local
gradients computed without communication (and equal on both ranks)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from mmf_b200.visual_bert import B200VisualBERT
from mmf_b200.ddp import B200DataParallel

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
cfg = types.SimpleNamespace(hidden_size=128, num_attention_heads=2, intermediate_size=256, num_hidden_layers=3, vocab_size=100,
                            max_position_embeddings=64, type_vocab_size=2, visual_embedding_dim=64, hidden_dropout_prob=0.0,
                            attention_probs_dropout_prob=0.0, layer_norm_eps=1e-12)
torch.manual_seed(0)
model = B200VisualBERT(cfg).to(dev).eval()
ddp = B200DataParallel(model, bucket_bytes=1 << 16)
g = torch.Generator(device=dev).manual_seed(100 + rank)
B, T, R = 4, 12, 9
sl = {"input_ids": torch.randint(0, 100, (B, T), device=dev, generator=g), "input_mask": torch.ones(B, T, dtype=torch.long, device=dev),
      "segment_ids": torch.zeros(B, T, dtype=torch.long, device=dev), "image_feature_0": torch.randn(B, R, 64, device=dev, generator=g).abs()}
w = torch.randn(B, T + R, 128, device=dev, generator=g)
# local gradients, no communication
with ddp.no_sync():
    model.zero_grad(set_to_none=True)
    (ddp(sl)["sequence_output"] * w).sum().backward()
local = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
# synchronised gradients
model.zero_grad(set_to_none=True)
(ddp(sl)["sequence_output"] * w).sum().backward()
torch.cuda.synchronize()
worst = 0.0
for n, p in model.named_parameters():
    if p.grad is None:
        continue
    mean = local[n].clone()
    dist.all_reduce(mean)
    mean /= world
    err = ((p.grad - mean).norm() / mean.norm().clamp_min(1e-6)).item()
    worst = max(worst, err)
    other = p.grad.detach().clone()
    dist.broadcast(other, src=0)
    assert torch.equal(other, p.grad), "rank %d differs from rank 0 on %s" % (rank, n)
assert worst < 1e-3, worst   # fp32 sum order / atomics only
if rank == 0:
    print("DDP check ok: world %d, worst rel diff vs mean of local grads %.2e, %d params" % (world, worst, len(local)))
dist.destroy_process_group()
