"""world_size=2 gloo test of the gradient-bucket logic in mmf_b200.ddp on CPU: the flat gradient buffers of the two
ranks end up equal to their mean, bucket slices cover the buffer exactly once, no_sync() skips communication.
(No kernels run: gradients are written into the flat buffer by hand, the way the weight-gradient GEMMs do on the GPU.)"""
import os
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, result):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mmf_b200.ddp import B200DataParallel
        from mmf_b200.engine import BertLayerW, ParamPack
        from mmf_b200.modules import B200BertEncoder
        torch.manual_seed(rank)   # different init per rank: the wrapper must broadcast rank 0's parameters
        cfg = types.SimpleNamespace(hidden_size=64, num_attention_heads=1, intermediate_size=128, num_hidden_layers=3,
                                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=1e-12)
        enc = B200BertEncoder(cfg)
        runner = enc._runner
        # build the pack on CPU by hand (ensure() insists on CUDA)
        params = []
        for m in runner.layers:
            params += BertLayerW.params(m)
        runner.pack = ParamPack(params, "cpu")
        per = len(BertLayerW.params(runner.layers[0]))
        runner.layer_param_ranges = [(runner.pack.offsets[i * per], 0) for i in range(3)]
        ddp = B200DataParallel(enc, bucket_bytes=1, overlap=False)
        p0 = torch.cat([p.detach().reshape(-1) for p in enc.parameters()])
        gathered = [torch.zeros_like(p0) for _ in range(world)]
        dist.all_gather(gathered, p0)
        assert torch.equal(gathered[0], gathered[1])                      # parameters were broadcast
        # emulate a backward: layer 2, 1, 0 complete in turn
        runner.pack.grad.copy_(torch.arange(runner.pack.total, dtype=torch.float32) * (rank + 1))
        sent = []
        orig = ddp._avg

        def spy(flat):
            sent.append((flat.data_ptr() - runner.pack.grad.data_ptr()) // 4)
            sent.append(flat.numel())
            orig(flat)
        ddp._avg = spy
        ddp._queue_finalize = lambda: None
        for i in (2, 1, 0):
            runner.grad_ready_hook(i)
        expect = torch.arange(runner.pack.total, dtype=torch.float32) * 1.5
        assert torch.allclose(runner.pack.grad, expect)
        offs, lens = sent[0::2], sent[1::2]
        assert sum(lens) == runner.pack.total and sorted(offs) == sorted(set(offs))   # every element sent exactly once
        # the same encoder behind a second autograd node of the same backward pass: its local gradients land on top
        # of the averaged ones and every region is sent again -> still the mean of the totals
        local2 = torch.arange(runner.pack.total, dtype=torch.float32).flip(0) * (rank + 1)
        runner.pack.on_reentry()
        runner.pack.grad.add_(local2)
        for i in (2, 1, 0):
            runner.grad_ready_hook(i)
        assert torch.allclose(runner.pack.grad, expect + torch.arange(runner.pack.total, dtype=torch.float32).flip(0) * 1.5)
        runner.pack.grad.copy_(expect)
        # no_sync: nothing is sent
        sent.clear()
        ddp._state.clear()
        with ddp.no_sync():
            for i in (2, 1, 0):
                runner.grad_ready_hook(i)
        assert not sent
        ddp.reduce_now()
        assert torch.allclose(runner.pack.grad, expect)
        result[rank] = True
    finally:
        dist.destroy_process_group()


def test_bucket_allreduce_world2():
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_worker, args=(2, port, result), nprocs=2, join=True)
    assert result.get(0) and result.get(1)


def _e2e_worker(rank, world, port, result):
    """Whole data-parallel step on CPU: the real host code (encoder module, autograd Function, bucket hooks, end-of-backward
    callback) over the kernel test double (tests/fake_kernels.py); both ranks must end with the average of the
    single-process gradients of their two batches."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_kernels as FK
        import mmf_b200.engine as E
        import mmf_b200.modules as M
        from mmf_b200.ddp import B200DataParallel
        E.F = FK
        M._require_cuda = lambda t, what: None
        cfg = types.SimpleNamespace(hidden_size=64, num_attention_heads=1, intermediate_size=128, num_hidden_layers=3,
                                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=1e-12)
        torch.manual_seed(100 + rank)                     # different init per rank: the wrapper broadcasts rank 0's
        enc = M.B200BertEncoder(cfg).eval()
        ddp = B200DataParallel(enc, bucket_bytes=1, overlap=False)
        g = torch.Generator().manual_seed(7)
        xs = [torch.randn(2, 6, 64, generator=g) for _ in range(world)]
        ws = [torch.randn(2, 6, 64, generator=g) for _ in range(world)]
        (ddp(xs[rank], None)[0] * ws[rank]).sum().backward()
        got = {k: p.grad.detach().clone() for k, p in enc.named_parameters()}
        # single-process reference: mean over the two batches on an un-wrapped copy with the same (rank 0) weights
        ref_enc = M.B200BertEncoder(cfg).eval()
        ref_enc.load_state_dict(enc.state_dict())
        acc = None
        for r in range(world):
            ref_enc.zero_grad(set_to_none=True)
            (ref_enc(xs[r], None)[0] * ws[r]).sum().backward()
            cur = {k: p.grad.detach().clone() for k, p in ref_enc.named_parameters()}
            acc = cur if acc is None else {k: acc[k] + cur[k] for k in acc}
        for k in got:
            ref = acc[k] / world
            err = (got[k] - ref).norm() / ref.norm().clamp_min(1e-3 * ref.numel() ** 0.5)
            assert err < 1e-3, (k, float(err))
        # second step under no_sync(): local gradients only, then reduce_now() gives the mean of the accumulated ones
        enc.zero_grad(set_to_none=True)
        with ddp.no_sync():
            (ddp(xs[rank], None)[0] * ws[rank]).sum().backward()
        local = enc.layer[0].output.dense.weight.grad.detach().clone()
        ddp.reduce_now()
        ref = acc["layer.0.output.dense.weight"] / world
        assert (enc.layer[0].output.dense.weight.grad - ref).norm() / ref.norm() < 1e-3
        assert (local - ref).norm() / ref.norm() > 1e-2           # before the reduction it really was rank-local
        # the encoder applied twice in one graph (two autograd nodes on one pack): the second node re-enters the pack,
        # the wrapper re-sends the regions, and the result is still the mean over ranks of the summed gradients
        enc.zero_grad(set_to_none=True)
        o = 1 - rank
        ((ddp(xs[rank], None)[0] * ws[rank]).sum() + (enc(xs[o], None)[0] * ws[o]).sum()).backward()
        k = "layer.1.intermediate.dense.weight"
        ref = acc[k]                                       # every rank saw both batches -> mean == sum of the two
        assert (enc.layer[1].intermediate.dense.weight.grad - ref).norm() / ref.norm() < 1e-3
        result[rank] = True
    finally:
        dist.destroy_process_group()


def test_data_parallel_step_end_to_end_world2():
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_e2e_worker, args=(2, port, result), nprocs=2, join=True)
    assert result.get(0) and result.get(1)


def _tied_bf16_worker(rank, world, port, result):
    """ADVICE r1 (ddp.py:157): a packed parameter whose .grad is NOT its slice of the flat buffer must still be averaged.
    (a) a weight of the pack is also used by a torch-side head (the tied MLM decoder case): autograd installs the
    head's gradient first and adds the flat view into that separate tensor; (b) a bf16 pack hands autograd cast copies."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_kernels as FK
        import mmf_b200.engine as E
        import mmf_b200.modules as M
        from mmf_b200.ddp import B200DataParallel
        E.F = FK
        M._require_cuda = lambda t, what: None
        cfg = types.SimpleNamespace(hidden_size=64, num_attention_heads=1, intermediate_size=128, num_hidden_layers=2,
                                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=1e-12)
        g = torch.Generator().manual_seed(11)
        xs = [torch.randn(2, 5, 64, generator=g) for _ in range(world)]
        ws = [torch.randn(2, 5, 64, generator=g) for _ in range(world)]
        zs = [torch.randn(3, 128, generator=g) for _ in range(world)]

        def loss_of(enc, r):
            tied = enc.layer[0].output.dense.weight                     # [64, 128], lives in the encoder's pack
            head = torch.nn.functional.linear(zs[r], tied).square().sum()
            return (enc(xs[r], None)[0] * ws[r]).sum() + head

        for dtype, tol in ((torch.float32, 1e-3), (torch.bfloat16, 3e-2)):
            torch.manual_seed(200 + rank)
            enc = M.B200BertEncoder(cfg).eval().to(dtype)
            ddp = B200DataParallel(enc, bucket_bytes=1, overlap=False)
            zs_ = zs
            if dtype == torch.bfloat16:
                zs_ = [z.to(dtype) for z in zs]

            def loss_dd(r):
                tied = enc.layer[0].output.dense.weight
                head = torch.nn.functional.linear(zs_[r], tied).float().square().sum()
                return (ddp(xs[r].to(dtype), None)[0].float() * ws[r]).sum() + head
            loss_dd(rank).backward()
            got = {k: p.grad.detach().float().clone() for k, p in enc.named_parameters()}
            ref_enc = M.B200BertEncoder(cfg).eval().to(dtype)
            ref_enc.load_state_dict(enc.state_dict())
            acc = None
            for r in range(world):
                ref_enc.zero_grad(set_to_none=True)
                tied = ref_enc.layer[0].output.dense.weight
                head = torch.nn.functional.linear(zs_[r], tied).float().square().sum()
                ((ref_enc(xs[r].to(dtype), None)[0].float() * ws[r]).sum() + head).backward()
                cur = {k: p.grad.detach().float().clone() for k, p in ref_enc.named_parameters()}
                acc = cur if acc is None else {k: acc[k] + cur[k] for k in acc}
            for k in got:
                ref = acc[k] / world
                err = (got[k] - ref).norm() / ref.norm().clamp_min(1e-3 * ref.numel() ** 0.5)
                assert err < tol, (str(dtype), k, float(err))
            # every rank holds the same gradients afterwards
            flat = torch.cat([v.reshape(-1) for v in got.values()])
            both = [torch.zeros_like(flat) for _ in range(world)]
            dist.all_gather(both, flat)
            assert torch.equal(both[0], both[1]), str(dtype)
        result[rank] = True
    finally:
        dist.destroy_process_group()


def test_data_parallel_tied_and_bf16_pack_world2():
    port = 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_tied_bf16_worker, args=(2, port, result), nprocs=2, join=True)
    assert result.get(0) and result.get(1)


def _end_mode_worker(rank, world, port, result, payload="fp32"):
    """mode="end": hooks send nothing, the end-of-backward callback all-reduces each flat buffer once
    (payload "bf16": through a bf16 copy of the buffer - both ranks end with the same values, 2^-9 from the fp32 mean)"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_kernels as FK
        import mmf_b200.engine as E
        import mmf_b200.modules as M
        from mmf_b200.ddp import B200DataParallel
        E.F = FK
        M._require_cuda = lambda t, what: None
        cfg = types.SimpleNamespace(hidden_size=64, num_attention_heads=1, intermediate_size=128, num_hidden_layers=3,
                                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=1e-12)
        torch.manual_seed(300 + rank)
        enc = M.B200BertEncoder(cfg).eval()
        ddp = B200DataParallel(enc, mode="end", overlap=False, payload=payload)
        calls = []
        orig = ddp._avg
        ddp._avg = lambda flat: (calls.append(flat.numel()), orig(flat))[1]
        g = torch.Generator().manual_seed(17)
        xs = [torch.randn(2, 6, 64, generator=g) for _ in range(world)]
        ws = [torch.randn(2, 6, 64, generator=g) for _ in range(world)]
        (ddp(xs[rank], None)[0] * ws[rank]).sum().backward()
        assert calls == [enc._runner.pack.total], calls                  # one collective for the whole flat buffer
        ref_enc = M.B200BertEncoder(cfg).eval()
        ref_enc.load_state_dict(enc.state_dict())
        acc = None
        for r in range(world):
            ref_enc.zero_grad(set_to_none=True)
            (ref_enc(xs[r], None)[0] * ws[r]).sum().backward()
            cur = {k: p.grad.detach().clone() for k, p in ref_enc.named_parameters()}
            acc = cur if acc is None else {k: acc[k] + cur[k] for k in acc}
        tol = 1e-3 if payload == "fp32" else 1e-2
        for k, p in enc.named_parameters():
            ref = acc[k] / world
            assert (p.grad - ref).norm() / ref.norm().clamp_min(1e-3 * ref.numel() ** 0.5) < tol, k
        if payload == "bf16":      # every rank holds the SAME reduced values (replicas stay in lock-step)
            flat = enc._runner.pack.grad.clone()
            other = flat.clone()
            dist.broadcast(other, src=0)
            assert torch.equal(flat, other)
        result[rank] = True
    finally:
        dist.destroy_process_group()


def test_data_parallel_end_mode_world2():
    port = 35500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_end_mode_worker, args=(2, port, result), nprocs=2, join=True)
    assert result.get(0) and result.get(1)


def test_data_parallel_bf16_payload_world2():
    port = 37500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_end_mode_worker, args=(2, port, result, "bf16"), nprocs=2, join=True)
    assert result.get(0) and result.get(1)
