"""Run under torchrun (one process per GPU): the macro matrix assembled from cells sharded over the
ranks (+ halo sum over NCCL) equals the one assembled on a single GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import coefficients as Cf  # noqa: E402
from hommx_b200 import LinearElasticityStratifiedHMM, PoissonHMM, mesh  # noqa: E402
from hommx_b200 import ufl as pufl  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name in ("poisson3d", "elasticity3d"):
    if name == "poisson3d":
        mk = lambda **kw: PoissonHMM(mesh.create_unit_cube(6, 5, 4), Cf.smooth_sin(pufl), lambda x: 1.0, mesh.create_unit_cube(4, 4, 4), 0.125, **kw)  # noqa: E731
    else:
        mk = lambda **kw: LinearElasticityStratifiedHMM(  # noqa: E731
            mesh.create_box((0, 0, 0), (1, 0.4, 0.1), (4, 3, 2)), Cf.hooke_fibre_3d(pufl), lambda x: pufl.as_vector([0.0, 0.0, -1.0]),
            mesh.create_unit_cube(4, 4, 4), 0.01, Cf.dtheta_rotation_3d(pufl), **kw)  # fmt: skip
    sharded = mk(device=local)
    sharded._assemble_stiffness()
    single = mk(device=local, shard=False)
    single._assemble_stiffness()
    a, b = sharded._A_values, single._A_values
    err = np.abs(a - b).max() / np.abs(b).max()
    n_shared = sharded._dev["halo"].n if world > 1 else 0
    print(f"[rank {rank}] {name}: world={world} cells {sharded._dev['lo']}..{sharded._dev['hi']} shared slots {n_shared} "
          f"of {len(a)}  max rel diff sharded vs single {err:.2e}", flush=True)  # fmt: skip
    ok = ok and err < 1e-13


def raw_nccl_comm():
    """An ncclComm_t made with the NCCL library torch has loaded (ctypes), as a host without torch would hold one."""
    import ctypes as C
    import glob

    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*")) + ["libnccl.so.2"]
    lib = C.CDLL(cands[0], mode=C.RTLD_GLOBAL)
    class UniqueId(C.Structure):  # ncclUniqueId is passed BY VALUE: a struct, not an array
        _fields_ = [("internal", C.c_byte * 128)]

    uid = UniqueId()
    if rank == 0:
        rc = lib.ncclGetUniqueId(C.byref(uid))
        assert rc == 0, f"ncclGetUniqueId -> {rc}"
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    uid = UniqueId.from_buffer_copy(bytes(t.cpu().tolist()))
    comm = C.c_void_p()
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    lib.ncclCommInitRank.restype = C.c_int
    rc = lib.ncclCommInitRank(C.byref(comm), world, uid, rank)
    assert rc == 0, f"ncclCommInitRank -> {rc}"
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    return lib, comm


# hmx_halo_sum_dev: the exchange through the C ABI with a raw ncclComm_t equals the torch.distributed all-reduce
if world > 1:
    lib, comm = raw_nccl_comm()
    halo = sharded._dev["halo"]
    vals = sharded._dev["vals"]
    torch.manual_seed(rank)
    vals.copy_(torch.rand_like(vals))
    a = vals.clone()
    halo.sum(a)  # pack + torch all-reduce + unpack
    b = vals.clone()
    sharded._solver.set_stream(torch.cuda.current_stream().cuda_stream)
    sharded._solver.halo_sum_dev(comm.value, b, halo.slots, halo.n, torch.zeros_like(halo.buf))
    torch.cuda.synchronize()
    same = bool(torch.equal(a, b))
    print(f"[rank {rank}] hmx_halo_sum_dev (raw ncclComm_t) == torch all-reduce on {halo.n} shared slots: {same}", flush=True)
    ok = ok and same
    lib.ncclCommDestroy(comm)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
