"""Multi-GPU check of the data-parallel train step (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/ddp_check.py

* the replicas stay bit-identical after several steps (the NCCL bucket all-reduce averages the gradients on every rank);
* the step replayed as a CUDA graph (all-reduces captured on their side stream) equals the eagerly enqueued step bit for bit;
* N ranks x batch b with averaged gradients equal ONE process that averages the per-shard gradients itself (per-replica BatchNorm
  statistics): rank 0 recomputes every shard's gradient single-GPU and compares.
Prints "DDP_CHECK OK" on rank 0.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import threading
    threading.Timer(240.0, lambda: os._exit(3)).start()  # never hang a GPU box
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from gan_inpainting_b200.lib.models import networks
    from gan_inpainting_b200.trainer import InpaintingTrainer
    from gan_inpainting_b200.data import synthetic_batch

    R, B, steps = 128, 4, 4
    batches = [tuple(t.to(dev) for t in synthetic_batch(B, R, seed=500 + 10 * i, rank=rank)) for i in range(steps)]

    def make(seed):
        torch.manual_seed(seed)  # rank-dependent init on purpose: the trainer must broadcast rank 0's weights
        G = networks.UnetGenerator(1, 1, 7, ngf=64, norm_layer=networks.get_norm_layer("batch"), use_dropout=False).to(dev).train()
        D = networks.PatchGANDiscriminator(sigmoid=False, imagedim=R).to(dev).train()
        return G, D

    results = {}
    for graph in (False, True):
        G, D = make(rank)
        tr = InpaintingTrainer("wgan", G, D, recon="l1", distributed=True, bucket_mb=8.0, graph=graph)
        if not graph:
            first_grads = None
        for i, (g, m) in enumerate(batches):
            tr.step(g, m, update_g=True)
            if i == 0 and not graph:
                first_grads = (tr.fG.grads.clone(), tr.fD.grads.clone())
        torch.cuda.synchronize()
        results[graph] = (tr.fG.params.clone(), tr.fD.params.clone(), tr.scalars.clone())
        if graph:
            assert len(tr._graphs) == 1
            tr.release_graphs()  # before destroy_process_group: captured collectives pin the NCCL communicator
    # graph == eager
    for a, b in zip(results[False], results[True]):
        assert torch.equal(a, b), "graph replay differs from the eager step"
    # replicas identical
    for t in results[True][:2]:
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, t), "replicas diverged"
    # all-reduced D gradient of step 1 == mean over ranks of the gradient each shard computes without communication
    G, D = make(0)
    tr1 = InpaintingTrainer("wgan", G, D, recon="l1", distributed=False)
    for fg in (tr1.fG, tr1.fD):
        dist.broadcast(fg.params, src=0)
    import gan_inpainting_b200._lib as L
    L.WEIGHT_EPOCH += 1
    tr1.step(batches[0][0], batches[0][1], update_g=False)
    local_d = tr1.fD.grads.clone()
    mean_d = local_d.clone()
    dist.all_reduce(mean_d)
    mean_d /= world
    err = float((mean_d - first_grads[1]).norm() / first_grads[1].norm())
    assert err < 1e-5, f"all-reduced D gradient != mean of the shard gradients ({err})"
    dist.barrier()
    if rank == 0:
        print("DDP_CHECK OK world=%d d_grad_err=%.2e" % (world, err), flush=True)
    dist.destroy_process_group()
    os._exit(0)


if __name__ == "__main__":
    main()
