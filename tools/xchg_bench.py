"""Cost of the sum over the ranks: kin40k-shape sweep step (exchange in the kernel's tail) and the stand-alone exchange kernel (M = 384 path),
per exchange mode.  Launch with torchrun; rank 0 prints.  usage: torchrun --nproc-per-node R tools/xchg_bench.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gaussianprocessnode_b200 import SGPContext

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(rank)
for M in (512, 384):
    X = rng.standard_normal((10000, 8)); y = np.sin(X[:, 0]); Z = np.random.default_rng(99).standard_normal((M, 8))
    base = SGPContext(local); base.set_kernel(1.0, np.full(8, 2.0)); base.set_inducing(Z); base.set_data(X, y)
    base.sweep_timed_flushed(5, 256); t1 = base.sweep_timed_flushed(20, 256)
    ctx = SGPContext(local)
    uid = [SGPContext.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    ctx.set_kernel(1.0, np.full(8, 2.0)); ctx.set_inducing(Z); ctx.set_data(X, y)
    ctx.sweep_timed_flushed(5, 256)
    torch.cuda.synchronize(); dist.barrier()
    tr = ctx.sweep_timed_flushed(20, 256)
    t = torch.tensor([tr[0], tr[1]], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("R=%d mode=%s M=%d: single-GPU step %.4f ms | sharded step %.4f ms (main kernel %.4f) -> exchange overhead %.1f us" % (
            world, os.environ.get("SGP_XCHG_MODE", "push"), M, t1[0], t[0].item(), t[1].item(), (t[0].item() - t1[0]) * 1e3))
    if M > 384:        # in-kernel clocks of the exchange stages (one more sweep with the debug counters on)
        ctx.sweep_debug_clocks(); ctx.sweep_timed(1); ctx.sweep_timed(1)
        r = ctx.sweep_debug_clocks(); r = r[r[:, 0] >= 0]
        q8, q9, q5 = r[r[:, 2] == 8], r[r[:, 2] == 9], r[r[:, 2] == 5]
        if rank == 0 and len(q8):
            print("   in-kernel clocks, mean / max over CTAs: phase 2 %.0f / %.0f | publish A %.0f / %.0f | wait A %.0f / %.0f | reduce-scatter %.0f / %.0f | publish B + wait B %.0f / %.0f | expand %.0f / %.0f | final barrier before phase 2 %.0f / %.0f" % (
                q9[:, 3].mean(), q9[:, 3].max(), q8[:, 0].mean(), q8[:, 0].max(), q8[:, 1].mean(), q8[:, 1].max(), q8[:, 3].mean(), q8[:, 3].max(),
                q9[:, 0].mean(), q9[:, 0].max(), q9[:, 1].mean(), q9[:, 1].max(), q5[:, 0].mean(), q5[:, 0].max()))
    ctx.close(); base.close()
dist.destroy_process_group()
