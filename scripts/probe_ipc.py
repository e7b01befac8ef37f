"""Probe: do CUDA IPC memory handles open across the ranks of one box (peer stores over NVLink)?  Run under torchrun."""
import os, sys
import torch, torch.distributed as td
from cuda import cudart
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
td.init_process_group('nccl')
err, ptr = cudart.cudaMalloc(1 << 20)
assert err == cudart.cudaError_t.cudaSuccess, err
cudart.cudaMemset(ptr, 0, 1 << 20)
err, h = cudart.cudaIpcGetMemHandle(ptr)
assert err == cudart.cudaError_t.cudaSuccess, err
hs = [None] * world
td.all_gather_object(hs, bytes(h.reserved))
peers = []
for r in range(world):
    if r == rank:
        peers.append(ptr); continue
    hh = cudart.cudaIpcMemHandle_t()
    hh.reserved = hs[r]
    err, p = cudart.cudaIpcOpenMemHandle(hh, cudart.cudaIpcMemLazyEnablePeerAccess)
    print('rank', rank, 'open handle of', r, '->', err)
    assert err == cudart.cudaError_t.cudaSuccess
    peers.append(p)
# every rank writes its rank+1 into word `rank` of every peer's buffer
src = torch.full((4,), rank + 1, dtype=torch.int32, device='cuda')
for r in range(world):
    (err,) = cudart.cudaMemcpy(peers[r] + 16 * rank, src.data_ptr(), 16, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
    assert err == cudart.cudaError_t.cudaSuccess, err
torch.cuda.synchronize(); td.barrier()
out = torch.empty(4 * world, dtype=torch.int32, device='cuda')
cudart.cudaMemcpy(out.data_ptr(), ptr, 16 * world, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
print('rank', rank, 'buffer', out.cpu().tolist())
