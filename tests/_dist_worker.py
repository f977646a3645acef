"""Worker for tests/test_distributed_cpu.py: one rank of a world-size-2 gloo job on the CPU."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200"), HERE]


def run(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import copy

    import torch
    import torch.distributed as dist

    import _cpu_backend
    from _helpers import fill_deterministic
    from lmnet_b200 import distributed as D
    from lmnet_b200.model import LM_Net
    from lmnet_b200.train import build_training, synthetic_batches, train_step

    _cpu_backend.apply(lambda obj, name, val: setattr(obj, name, val))
    torch.set_num_threads(2)
    args = D.init_distributed_mode()
    assert args.distributed and args.dist_backend == "gloo" and D.get_world_size() == world and D.get_rank() == rank
    assert D.is_main_process() == (rank == 0)

    net = LM_Net(3, 2)
    fill_deterministic(net, seed=1)            # identical weights on every rank
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    local = copy.deepcopy(net)                 # un-wrapped twin for the manual all-reduce check
    ddp = D.wrap_ddp(net, "cpu")
    assert ddp is not net
    images, labels = synthetic_batches(1, 1, 32, seed=100 + rank, pin=False)[0]   # rank-specific shard

    crit = torch.nn.CrossEntropyLoss()
    out = ddp(images)
    torch.nn.functional.cross_entropy(out, labels).backward()
    out2 = local(images)
    torch.nn.functional.cross_entropy(out2, labels).backward()
    worst, scale = 0.0, 0.0
    for (n, p), (_, q) in zip(net.named_parameters(), local.named_parameters()):
        g = q.grad.clone()
        D.reduce_value(g, average=True)        # reference helper: all-reduce + divide by world size
        scale = max(scale, float(g.abs().max()))
        worst = max(worst, float((p.grad - g).abs().max()))
    worst /= scale                             # biases in front of a BatchNorm have ~0 gradient: use a global scale
    # one optimiser step through the harness, then parameters must agree across ranks
    opt, c, d = build_training(ddp, "cpu", fused=False)
    train_step(ddp, opt, images, labels, c, d, amp_dtype=None)
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    # BatchNorm buffers are per process (reference: no SyncBN): running means differ between shards
    rm = net.conv1[0].large_conv.bn.running_mean.clone()
    rms = [torch.empty_like(rm) for _ in range(world)]
    dist.all_gather(rms, rm)
    # the flat-buffer data-parallel step used with CUDA graphs (eager here: no GPU) must give the same update
    from lmnet_b200.train import GraphedTrainStep

    torch.manual_seed(7 + rank)                # different init per rank: the constructor must broadcast rank 0's
    flat_net = LM_Net(3, 2)
    for m in flat_net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    ref_net = LM_Net(3, 2)
    for m in ref_net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    opt_f, c_f, d_f = build_training(flat_net, "cpu", fused=False)
    step = GraphedTrainStep(flat_net, opt_f, c_f, d_f, images, labels, amp_dtype=None)
    assert step.graph is None and step.flat is not None
    ref_net.load_state_dict(flat_net.state_dict())          # after the broadcast
    ref_ddp = D.wrap_ddp(ref_net, "cpu")
    opt_r, c_r, d_r = build_training(ref_ddp, "cpu", fused=False)
    step(images, labels)
    train_step(ref_ddp, opt_r, images, labels, c_r, d_r, amp_dtype=None)
    # compare the averaged gradients (AdamW's first update is +-lr whatever the gradient's size, so parameters of
    # ~zero-gradient entries may legitimately differ by 2*lr)
    gscale = max(float(b.grad.abs().max()) for b in ref_net.parameters())
    flat_vs_ddp = max(float((a.grad - b.grad).abs().max()) for a, b in zip(flat_net.parameters(), ref_net.parameters())) / gscale
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(f"{worst} {int(same)} {int(not torch.equal(rms[0], rms[1]))} {flat_vs_ddp}\n")
    D.cleanup()


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
