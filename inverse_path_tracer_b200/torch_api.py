"""torch front-end of the device-pointer entry points (include/ipt_b200.h part 2).

PyTorch is plumbing here: it owns device memory, the CUDA stream and torch.distributed; every kernel
is launched by libipt_b200.so on torch's current stream, so torch.cuda.Event timing and stream
ordering with NCCL collectives work as usual.  This is the "PyTorch extension for the Python-facing
inverse loop": RenderLoss is a torch.autograd.Function whose backward is the path adjoint.
"""
import ctypes as C

import torch

from . import native


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dp(t):
    return C.c_void_p(t.data_ptr())


class TorchScene:
    """A scene resident on one GPU, rendering into torch tensors."""

    def __init__(self, objects, asset_root=None, device=None, traversal_mode=0):
        if not torch.cuda.is_available():
            raise RuntimeError("inverse_path_tracer_b200 needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.scene = native.Scene(objects, asset_root=asset_root, device=self.device.index, traversal_mode=traversal_mode)
        self.nT = self.scene.nT

    @classmethod
    def from_file(cls, scenefile, asset_root=None, **kw):
        from .ipt_cuda import parse_scene_file
        return cls(parse_scene_file(scenefile), asset_root=asset_root, **kw)

    def close(self):
        self.scene.close()

    # ---- materials ------------------------------------------------------------------------------
    def materials(self):
        return torch.from_numpy(self.scene.get_materials()).to(self.device)

    def set_materials(self, kd):
        kd = kd.detach().to(self.device, torch.float32).contiguous()
        assert kd.numel() == 3 * self.nT
        native._check(native.lib().iptSetMaterialsDevice(self.scene.h, _dp(kd), _stream_ptr()), "iptSetMaterialsDevice")
        self._kd_keepalive = kd

    # ---- forward --------------------------------------------------------------------------------
    def render(self, p, out=None, out_u8=None, counters=None):
        rows = native.Scene._rows(p)
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty((rows, p.width, 3), dtype=torch.float32, device=self.device)
            native._check(native.lib().iptRenderDevice(self.scene.h, C.byref(p), _dp(out), _dp(out_u8) if out_u8 is not None else None,
                                                       _dp(counters) if counters is not None else None, _stream_ptr()), "iptRenderDevice")
        return out

    def trace(self, rays):
        rays = rays.to(self.device, torch.float32).contiguous().view(-1, 6)
        n = rays.shape[0]
        tri = torch.empty(n, dtype=torch.int32, device=self.device)
        t = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            native._check(native.lib().iptTraceRaysDevice(self.scene.h, _dp(rays), n, _dp(tri), _dp(t), _stream_ptr()), "iptTraceRaysDevice")
        return tri, t

    # ---- inverse --------------------------------------------------------------------------------
    def graph_raw(self, p, target_u8):
        """raw Edge sums of this row shard: ((nT+1)*nT, 16) float64 on the device (all-reduce, then native.graph_compress)"""
        assert target_u8.dtype == torch.uint8 and target_u8.is_cuda and target_u8.is_contiguous()
        raw = torch.empty(((self.nT + 1) * self.nT, 16), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            native._check(native.lib().iptGraphAccumulateDevice(self.scene.h, C.byref(p), _dp(target_u8), _dp(raw), _stream_ptr()),
                          "iptGraphAccumulateDevice")
        return raw

    def graph_data(self, p, target_u8, group=None, reduce=False):
        """createGraph's data product on the device: (w [(nT+1), nT], pixel [(nT+1), nT, 3], light [(nT+1), nT, 3]), the
        layout of ipt_cuda.generate_data (eye row last).  reduce=True sums the raw edge statistics over the row shards
        of `group` first (DataWrapper::compress is not linear, so it runs after the all-reduce)."""
        raw = self.graph_raw(p, target_u8)
        if reduce and torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(raw, group=group)
        n = (self.nT + 1) * self.nT
        data = torch.empty(7 * n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            native._check(native.lib().iptGraphCompressDevice(_dp(raw), self.nT, _dp(data), _stream_ptr()), "iptGraphCompressDevice")
        return (data[:n].view(self.nT + 1, self.nT), data[n:4 * n].view(self.nT + 1, self.nT, 3), data[4 * n:].view(self.nT + 1, self.nT, 3))

    def loss_and_grad(self, p, target_u8, out=None, img=None):
        """[loss, dLoss/dKd (nT*3)] float64 on the device for this row shard (sum over shards = whole image)"""
        assert target_u8.dtype == torch.uint8 and target_u8.is_cuda and target_u8.is_contiguous()
        if out is None:
            out = torch.empty(1 + 3 * self.nT, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            native._check(native.lib().iptRenderGradDevice(self.scene.h, C.byref(p), _dp(target_u8), _dp(out),
                                                           _dp(img) if img is not None else None, _stream_ptr()), "iptRenderGradDevice")
        return out


class RenderLoss(torch.autograd.Function):
    """loss(kd) = mean over pixels/channels of (L/(1+L) - target/255)^2, differentiable in the albedos.

    forward runs the forward render + the path adjoint in one call (the adjoint replays the same
    Philox streams); `group` (optional torch.distributed group) sums loss and gradient over row shards.
    """

    @staticmethod
    def forward(ctx, kd, tscene, params, target_u8, group=None, reduce=True):
        tscene.set_materials(kd)
        lg = tscene.loss_and_grad(params, target_u8)
        if reduce and (group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())):
            torch.distributed.all_reduce(lg, group=group)
        ctx.save_for_backward(lg[1:].view(-1, 3).to(kd.dtype))
        return lg[0].to(kd.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        return grad_out * g, None, None, None, None, None


def render_loss(kd, tscene, params, target_u8, group=None, reduce=True):
    """reduce=True: `params` describes this rank's row shard of ONE scene, loss and gradient are summed over the
    ranks of `group`.  reduce=False: the scene belongs to this rank alone (scene-sharded batches, BASELINE config 4)."""
    return RenderLoss.apply(kd, tscene, params, target_u8, group, reduce)


def inverse_optimizer(tscene, params, target_u8, group=None, **adam):
    """native.InverseOptimizer for `tscene` (BASELINE config 3 / 4): the whole iteration -- render, loss, adjoint, sum
    over the ranks of `group`, Adam, albedo update -- runs inside the library, enqueued by one call.  `params` describes
    this rank's row shard; with an initialised process group of more than one rank the mailboxes are exchanged once
    (CUDA IPC handles through all_gather_object) and every later step talks GPU to GPU over NVLink."""
    assert target_u8.dtype == torch.uint8 and target_u8.is_cuda and target_u8.is_contiguous()
    rank, world = 0, 1
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank, world = torch.distributed.get_rank(group), torch.distributed.get_world_size(group)
    opt = native.InverseOptimizer(tscene.scene, params, target_u8.data_ptr(), rank=rank, world=world, **adam)
    opt._target_keepalive = target_u8
    if world > 1:
        handles = [None] * world
        torch.distributed.all_gather_object(handles, opt.ipc_handle(), group=group)
        opt.connect_ipc(handles)
        torch.distributed.barrier(group=group)
    return opt
