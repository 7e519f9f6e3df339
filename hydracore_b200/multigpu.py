"""Multi-GPU plumbing: one process per GPU (torch.distributed), replicated scene, image plane partitioned in interleaved tiles
(PT / MISPT) or sample indices partitioned modulo the world size (MISPT-QMC), HDR SUM buffers combined over NVLink.

The reference's multi-GPU mode is one OS process per GPU adding partial framebuffers into a shared-memory image under a mutex
(GPUOCLLayerOther.cpp:365-430, README.md:99-103).  Here the exchange lives in the product library (hc_comm.cu: hc_comm_init / hc_fb_reduce,
NCCL send/recv of the owned tiles or ncclReduce of full-size buffers); this module only carries the NCCL unique id from rank 0 to the
other ranks through torch.distributed (join_communicator) and mirrors the ownership rules for the CPU tests (gloo).  There is no
per-bounce exchange, hence no collective inside any kernel."""
import numpy as np


def tile_owner_map(width, height, tile, world):
    """rank that owns each pixel: tile t (row-major over ceil(W/T) x ceil(H/T)) -> rank t mod world.  Mirrors BuildOwnedPixels (hc_path.cu)."""
    tx = (width + tile - 1)//tile
    ys, xs = np.mgrid[0:height, 0:width]
    return (((ys//tile)*tx + (xs//tile)) % world).astype(np.int32)


def owned_pixels(width, height, tile, rank, world):
    """Pixel indices (y*W + x) of `rank`, in the tile-by-tile order the device uses."""
    T = max(1, tile)
    tx, ty = (width + T - 1)//T, (height + T - 1)//T
    out = []
    for t in range(tx*ty):
        if t % world != rank:
            continue
        x0, y0 = (t % tx)*T, (t//tx)*T
        x1, y1 = min(width, x0 + T), min(height, y0 + T)
        for by in range(y0, y1, 4):                 # 8 x 4 pixel blocks inside the tile (one warp = one compact screen patch)
            for bx in range(x0, x1, 8):
                yy, xx = np.mgrid[by:min(y1, by + 4), bx:min(x1, bx + 8)]
                out.append((yy*width + xx).reshape(-1))
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def qmc_sample_range(width, height, rank, world):
    """Sample slots of one QMC pass owned by `rank`: i = k*world + rank < W*H  (k_pt_generate, hc_path.cu)."""
    return np.arange(rank, width*height, world, dtype=np.int64)


def join_communicator(layer, dist, device=None):
    """Create the library's NCCL communicator across the ranks of `dist`: rank 0 makes the unique id (hc_comm_unique_id), torch.distributed
    broadcasts its 128 bytes (plumbing only), every rank calls hc_comm_init."""
    import torch
    from .layer import CudaLayer
    if dist is None or dist.get_world_size() == 1:
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    t = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(CudaLayer.CommUniqueId()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    layer.CommInit(bytes(t.cpu().numpy().tobytes()), rank, world)


def reduce_sums(dist, tensor, dst=0):
    """Combine per-rank SUM framebuffers (float32, 4*W*H) on `dst`.  Tile partitions are disjoint, so SUM is also a gather."""
    if dist is None or dist.get_world_size() == 1:
        return tensor
    dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM)
    return tensor


class DevicePointerTensor:
    """Expose a raw device pointer (hc_fb_device_ptr) to torch through __cuda_array_interface__ without copying."""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {"shape": (int(nfloats),), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}


def framebuffer_tensor(layer, device):
    """torch view of the layer's per-pixel SUM buffer (float32[4*W*H]) living in the library's allocation."""
    import torch
    ptr, n = layer.fb_device_ptr()
    return torch.as_tensor(DevicePointerTensor(ptr, n), device=device)


def reduce_framebuffer(layer, dist, device, dst=0):
    """FinishAll on the library's stream, then NCCL reduce(SUM) of the SUM buffer in place; returns the tensor (complete on `dst`)."""
    import torch
    layer.FinishAll()
    t = framebuffer_tensor(layer, device)
    reduce_sums(dist, t, dst)
    torch.cuda.synchronize(device)
    return t
