"""Multi-GPU check of the data-parallel train step AGAINST THE ORACLE (run under torchrun, one rank per GPU; TEST INFRASTRUCTURE —
it imports oracle/, which tools/ may not):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/ddp_oracle_check.py

On top of tools/ddp_check.py's self-consistency checks (replicas identical, graph replay == eager step, all-reduced gradient ==
mean of the shard gradients) rank 0 restates what N-rank data parallelism with per-replica BatchNorm statistics means in terms of
the reference ("N reference replicas at batch b with averaged gradients"): oracle.steps.wgan_step
(experiment_list/wgan_l1.py:110-186) is run on every rank's shard from the same weights, the D gradients are averaged and stepped
(RMSprop + clip), every shard's G step continues from those D weights, the G gradients are averaged and stepped.  The N-rank CUDA
result is compared with that, free-running (tests/test_steps_gpu.py tier B bounds).  Prints "DDP_ORACLE_CHECK OK" on rank 0.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def oracle_ddp_step(nets, steps, gp, dp, shards, clip=0.01):
    """N oracle replicas, averaged gradients.  Returns (G params, D params, mean d_grads, mean g_grads) after one wgan_l1 step."""
    cl = lambda p: steps.enable_grad({k: v.clone() for k, v in p.items()})
    n = len(shards)
    # ---- D step on every shard (the G update is skipped: only the D gradients are wanted here)
    d_grads = []
    for ground, mask in shards:
        G_, D_ = cl(gp), cl(dp)
        o = steps.wgan_step(G_, D_, steps.make_optimizer("rmsprop", G_), steps.make_optimizer("rmsprop", D_), ground, mask, update_g=False)
        d_grads.append(o["d_grads"])
    mean_d = {k: sum(g[k] for g in d_grads) / n for k in d_grads[0] if d_grads[0][k] is not None}
    D_new = cl(dp)
    opt = steps.make_optimizer("rmsprop", D_new)
    for k, v in nets.trainable(D_new).items():
        v.grad = mean_d[k].clone()
    opt.step()
    with torch.no_grad():
        for v in steps.leaves(D_new):
            v.clamp_(-clip, clip)
    d_after = {k: v.detach().clone() for k, v in D_new.items()}

    def adopt(params):  # after_d_step hook: every replica continues with the all-reduced D update
        with torch.no_grad():
            for k, v in params.items():
                if v.dtype.is_floating_point and "running" not in k:
                    v.copy_(d_after[k])

    # ---- G step on every shard from the averaged D
    g_grads = []
    for ground, mask in shards:
        G_, D_ = cl(gp), cl(dp)
        o = steps.wgan_step(G_, D_, steps.make_optimizer("rmsprop", G_), steps.make_optimizer("rmsprop", D_), ground, mask, after_d_step=adopt)
        g_grads.append(o["g_grads"])
    mean_g = {k: sum(g[k] for g in g_grads) / n for k in g_grads[0]}
    G_new = cl(gp)
    opt = steps.make_optimizer("rmsprop", G_new)
    for k, v in nets.trainable(G_new).items():
        v.grad = mean_g[k].clone()
    opt.step()
    return G_new, D_new, mean_d, mean_g


def main():
    import threading
    threading.Timer(400.0, lambda: os._exit(3)).start()  # never hang a GPU box
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from gan_inpainting_b200.trainer import InpaintingTrainer
    from oracle import nets, steps, synth
    from tests.test_networks_gpu import make_D, make_G
    from tests.util import cos, rel

    R, B = 128, 4
    shards = [synth.make_batch(B, R, seed=900 + r, soft_edge=True) for r in range(world)]
    gp, dp = nets.init_unet(0), nets.init_patchgan(1, res=R)
    G, D = make_G(dev, gp).train(), make_D(dev, dp, R, False).train()
    tr = InpaintingTrainer("wgan", G, D, recon="l1", distributed=True, bucket_mb=8.0, graph=False)
    ground, mask = shards[rank]
    tr.step(ground.to(dev), mask.to(dev), update_g=True)
    torch.cuda.synchronize()
    # replicas identical after the step
    for t in (tr.fG.params, tr.fD.params):
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, t), "replicas diverged"
    ok = torch.ones(1, device=dev)
    if rank == 0:
        try:
            G_o, D_o, mean_d, mean_g = oracle_ddp_step(nets, steps, gp, dp, shards)
            worst_cos = 1.0
            for name, g in zip(tr.fG.names, tr.fG.grad_views):  # the all-reduced (averaged) gradients the optimizer consumed
                worst_cos = min(worst_cos, cos(g, mean_g[name]))
            for name, g in zip(tr.fD.names, tr.fD.grad_views):
                if float(mean_d[name].norm()) > 1e-12:
                    worst_cos = min(worst_cos, cos(g, mean_d[name]))
            assert worst_cos > 0.9, f"averaged gradients vs the N-replica oracle: cos {worst_cos}"
            # post-step weights: the first RMSprop step is sign descent of 5e-4 per element (tests/test_steps_gpu.py: check_params)
            worst = 0.0
            for net, ref in ((G, G_o), (D, D_o)):
                for k, v in net.state_dict().items():
                    if v.dtype.is_floating_point and "running" not in k:
                        r_ = ref[k].detach().float()
                        err = float((v.detach().float().cpu() - r_).norm())
                        bound = 1e-2 * float(r_.norm()) + 2 * 5e-4 * (0.15 * r_.numel()) ** 0.5
                        assert err <= bound, (k, err, bound)
                        worst = max(worst, err / max(bound, 1e-30))
            print("DDP_ORACLE_CHECK OK world=%d worst gradient cosine %.4f, worst weight error / bound %.3f" % (world, worst_cos, worst), flush=True)
        except AssertionError as e:
            print("DDP_ORACLE_CHECK FAILED", e, flush=True)
            ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0 if float(ok) > 0 else 1)


if __name__ == "__main__":
    main()
