"""Worker of tests/test_multi_gpu.py: one process per GPU under torch.distributed.run.  Every rank decodes its shard of
one batch; the counters are summed over the ranks by gswm.Comm (mailboxes mapped over NVLink with CUDA IPC) -- stand-alone
and fused into the extract kernel -- and must equal the single-GPU totals and the NCCL sum.  Prints 'MGPU OK' on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a-watermark-for-diffusion-models_b200"))
import gswm  # noqa: E402


def _own_nccl_comm(rank, world, dev):
    """An ncclComm_t created with ctypes (ncclGetUniqueId on rank 0, broadcast through torch.distributed, ncclCommInitRank):
    what a C host that owns a communicator would hand to gswm_allreduce_counters.  None if libnccl cannot be loaded."""
    import ctypes as C
    import glob
    import site

    cands = []
    for sp in site.getsitepackages() + [os.path.dirname(os.path.dirname(torch.__file__))]:
        cands += glob.glob(os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so*"))
    lib = None
    for path in cands + ["libnccl.so.2", "libnccl.so"]:
        try:
            lib = C.CDLL(path)
            break
        except OSError:
            continue
    if lib is None:
        return None

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_byte * 128)]

    uid = UniqueId()
    if rank == 0:
        assert lib.ncclGetUniqueId(C.byref(uid)) == 0
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    C.memmove(C.byref(uid), bytes(t.cpu().tolist()), 128)
    comm = C.c_void_p()
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    assert lib.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    return lib, comm


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = gswm.Comm(dev)
    assert comm.world == world and comm.rank == rank
    total, shape, L = 1003, (4, 64, 64), 256                      # 1003: ragged shards
    km = gswm.KeyMaterial.make(bytes.fromhex(gswm.DEFAULT_KEY_HEX), bytes.fromhex(gswm.DEFAULT_NONCE_HEX),
                               gswm.pad_message("lthero", L // 8), L)
    lo, hi = gswm.sharding.shard_range(total, rank, world)
    z = gswm.embed_batch(hi - lo, shape, km, 0x5EED, 0, lo, dev)
    g = torch.Generator(dev).manual_seed(1234)                    # same noise stream on every rank, sliced by global index
    noise = torch.randn((total, *shape), device=dev, generator=g)[lo:hi]
    zn = z + 4.3 * noise                                          # ~90 % decoded-bit accuracy: non-trivial counters
    # the single-GPU answer, computed redundantly on every rank
    z_all = gswm.embed_batch(total, shape, km, 0x5EED, 0, 0, dev)
    assert torch.equal(z_all[lo:hi], z)                           # sharding does not change the latents
    want = gswm.extract_batch(z_all + 4.3 * torch.randn((total, *shape), device=dev, generator=torch.Generator(dev).manual_seed(1234)),
                              km).counters.clone()
    # 1. stand-alone all-reduce, several epochs
    for _ in range(5):
        res = gswm.extract_batch(zn, km)
        got = gswm.sharding.allreduce_counters(res.counters.clone(), comm=comm)
        assert torch.equal(got, want), (rank, got.tolist(), want.tolist())
    # 2. fused into the extract kernel
    for _ in range(5):
        res = gswm.extract_batch(zn, km, comm=comm)
        assert torch.equal(res.reduced, want), (rank, res.reduced.tolist(), want.tolist())
    # 3. the NCCL form: c10d's all-reduce, and gswm_allreduce_counters on a communicator of our own (torch's ncclComm_t is
    #    not reachable from Python, so one is created through ctypes on the libnccl torch itself has loaded)
    nc = gswm.extract_batch(zn, km).counters.clone()
    dist.all_reduce(nc)
    assert torch.equal(nc, want)
    nccl = _own_nccl_comm(rank, world, dev)
    if nccl is not None:
        lib, comm_ptr = nccl
        mine = gswm.extract_batch(zn, km).counters.clone()
        st = torch.cuda.current_stream(dev).cuda_stream
        rc = gswm._lib.lib().gswm_allreduce_counters(comm_ptr, mine.data_ptr(), mine.numel(), st)
        torch.cuda.synchronize(dev)
        assert rc == 0 and torch.equal(mine, want), (rank, rc, mine.tolist())
        assert gswm._lib.lib().gswm_allreduce_counters(None, mine.data_ptr(), 6, st) == -1          # GSWM_E_NULL
        lib.ncclCommDestroy(comm_ptr)
    elif rank == 0:
        print("MGPU note: libnccl not loadable through ctypes, gswm_allreduce_counters not exercised", flush=True)
    # 4. ranks arriving at very different times (rank 0 late): the exchange waits, nothing is lost
    if rank == 0:
        torch.cuda._sleep(int(2e8))
    res = gswm.extract_batch(zn, km, comm=comm)
    assert torch.equal(res.reduced, want)
    # 5. fewer latents than ranks: the last rank's shard is empty, it decodes nothing and still joins the exchange
    small = world - 1
    lo2, hi2 = gswm.sharding.shard_range(small, rank, world)
    z2 = gswm.embed_batch(hi2 - lo2, shape, km, 7, 0, lo2, dev)
    r2 = gswm.extract_batch(z2, km, comm=comm)
    assert r2.reduced.tolist() == [small * L, small * L, small, small, 0, 0], (rank, r2.reduced.tolist())
    assert (hi2 - lo2 == 0) == (rank == world - 1)
    torch.cuda.synchronize()
    assert comm.status() == 0
    assert 0.85 < float(want[0]) / float(want[1]) < 0.95 and int(want[3]) == total
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU OK", world, want.tolist(), "nccl_c_abi=%d" % int(nccl is not None), flush=True)


if __name__ == "__main__":
    main()
