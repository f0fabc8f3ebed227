"""Whole-frame rendering on 1..N GPUs of one box — the replacement of the reference's per-frame driver,
the swarm lambda src/main.cpp:139-154 (16 CPU threads in 4x4 screen tiles) + samples_to_image (:156-158).

One process per GPU.  The voxel scene is replicated; the frame's 4-row tiles are dealt round-robin to the
ranks (sky and terrain rows cost very differently, so contiguous slabs would be unbalanced); every rank
renders and resolves its tiles with libvrt's kernels and the RGBA slabs are exchanged with ONE NCCL
all-gather over NVLink.  Because the RNG is counter based (pixel, sample, dimension) the assembled frame
is bit-identical for every world size.

torch is used for device buffers, the stream and torch.distributed only.
"""
import ctypes as C

import numpy as np
import torch

from . import capi
from .capi import check, lib, ptr


class TileExchange:
    """Round-robin 4-row-tile partition of a W x H RGBA frame over `world` ranks and its all-gather.
    Pure torch (any device / backend): NCCL on the GPUs, gloo in the CPU tests."""

    def __init__(self, width, height, rank, world, device, group=None):
        self.W, self.H, self.rank, self.world, self.group = int(width), int(height), int(rank), int(world), group
        unit = 4 * self.world
        self.H_pad = (self.H + unit - 1) // unit * unit
        self.tiles_local = self.H_pad // unit
        self.row_bytes = self.W * 4
        self.tile_bytes = 4 * self.row_bytes
        if self.world > 1:
            self.slab = torch.empty(self.tiles_local * self.tile_bytes, dtype=torch.uint8, device=device)
            self.gathered = torch.empty(self.world * self.slab.numel(), dtype=torch.uint8, device=device)
            self.frame = torch.empty(self.H_pad * self.row_bytes, dtype=torch.uint8, device=device)

    def owned_rows(self):
        """Rows this rank renders: 4-row tile t belongs to rank t % world."""
        return [y for y in range(self.H) if (y >> 2) % self.world == self.rank]

    def gather(self, rgba):
        """rgba: flat uint8 [H_pad*W*4] with this rank's rows filled → the assembled frame on every rank."""
        if self.world == 1:
            return rgba
        import torch.distributed as dist
        tiles = rgba.view(self.tiles_local, self.world, self.tile_bytes)                # [local tile, owner, bytes]
        self.slab.view(self.tiles_local, self.tile_bytes).copy_(tiles[:, self.rank, :])
        dist.all_gather_into_tensor(self.gathered, self.slab, group=self.group)
        g = self.gathered.view(self.world, self.tiles_local, self.tile_bytes)
        self.frame.view(self.tiles_local, self.world, self.tile_bytes).copy_(g.permute(1, 0, 2))
        return self.frame


class FrameRenderer:
    def __init__(self, scene, width, height, rank=0, world=1, group=None, device=None, stream=None):
        self.scene, self.W, self.H = scene, int(width), int(height)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.device = device if device is not None else torch.device("cuda", scene.ctx.device)
        # libvrt's launches and torch's copies/collectives must share one stream
        self.stream = stream if stream is not None else torch.cuda.Stream(self.device)
        scene.ctx.set_stream(self.stream.cuda_stream)      # libvrt re-applies an installed L2 access-policy window to it
        scene.ctx.torch_stream = self.stream               # keeps the cudaStream_t alive as long as the context uses it
        self.exchange = TileExchange(self.W, self.H, self.rank, self.world, self.device, group)
        self.H_pad, self.row_bytes = self.exchange.H_pad, self.exchange.row_bytes
        self.accum = torch.zeros(self.H_pad * self.W * 4, dtype=torch.int32, device=self.device)   # r,g,b,count sums
        self.rgba = torch.zeros(self.H_pad * self.W * 4, dtype=torch.uint8, device=self.device)
        self.host_frame = torch.empty(self.H * self.row_bytes, dtype=torch.uint8).pin_memory()
        self.use_gi, self.gi_bounces, self.use_samples = False, 1, True
        self.roughness, self.max_bounds = 0.0, 4
        self.checker_board_offset, self.checker_area_height = None, 0     # main.cpp:137,143 / :132
        self.display = None                                               # denoised_tex of main.cpp:159-177
        self.autofocus = False                                            # device-side centre-ray focus, main.cpp:114-121
        self.seed = (0x5EED, 0)
        self.light = np.zeros(3, np.float32)

    def params(self, spp, sample_offset=0):
        p = capi.RenderParams()
        p.width, p.height, p.row_begin, p.row_end = self.W, self.H, 0, self.H
        p.spp, p.sample_offset = int(spp), int(sample_offset)
        p.seed_lo, p.seed_hi = self.seed
        p.light_position[:] = [float(x) for x in self.light]
        p.use_gi, p.gi_bounces, p.use_samples = int(self.use_gi), int(self.gi_bounces), int(self.use_samples)
        p.tile_step, p.tile_index = self.world, self.rank
        p.roughness, p.max_bounds = float(self.roughness), int(self.max_bounds)
        p.checker = 0 if self.checker_board_offset is None else 1 + (int(self.checker_board_offset) & 1)
        p.checker_area_height = int(self.checker_area_height)
        p.autofocus = int(bool(self.autofocus))
        return p

    # -- device-resident frame: enqueue only (the caller owns stream/synchronisation) --
    def accumulate(self, cam_struct, p):
        check(lib().vrt_render_accumulate_device(self.scene.handle, C.byref(cam_struct), C.byref(p), ptr(self.accum)))

    def resolve(self, p):
        check(lib().vrt_render_resolve_device(self.scene.handle, C.byref(p), ptr(self.accum), ptr(self.rgba)))

    def gather(self):
        """All-gather of the ranks' RGBA tiles (the frame's one exchange step).  No-op on one GPU."""
        return self.exchange.gather(self.rgba)

    def render_device(self, camera, spp, sample_offset=0, clear=True):
        """One frame, device resident on every rank: clear, accumulate, resolve, gather.  Returns a uint8 view
        [H, W, 4] of the assembled frame."""
        p = self.params(spp, sample_offset)
        with torch.cuda.stream(self.stream):
            if clear:
                self.accum.zero_()
            self.accumulate(camera.as_struct(), p)
            self.resolve(p)
            frame = self.gather()
        return frame.view(self.H_pad, self.W, 4)[: self.H]

    def present_device(self, frame, median=0, old_value_conservation=0.1):
        """main.cpp:159-177 on the device: optional median, then the persistence blend of `frame` (a device uint8
        [H, W, 4] view such as render_device returns) into the display surface, which is returned."""
        if self.display is None:
            self.display = torch.zeros(self.H * self.W * 4, dtype=torch.uint8, device=self.device)
        p = capi.PresentParams(self.W, self.H, int(median), float(old_value_conservation))
        with torch.cuda.stream(self.stream):
            check(lib().vrt_present_device(self.scene.ctx.handle, ptr(frame), ptr(self.display), C.byref(p)))
        return self.display.view(self.H, self.W, 4)

    def render(self, camera, spp, sample_offset=0):
        """The user-facing call: a finished frame in host memory (pinned), as numpy [H, W, 4] uint8."""
        frame = self.render_device(camera, spp, sample_offset)
        with torch.cuda.stream(self.stream):
            self.host_frame.view(self.H, self.W, 4).copy_(frame, non_blocking=True)
        self.stream.synchronize()
        return self.host_frame.view(self.H, self.W, 4).numpy()

    def stats(self):
        st = capi.RenderStats()
        check(lib().vrt_scene_last_render_stats(self.scene.handle, C.byref(st)))
        return dict(rays=list(st.rays), complexity=list(st.complexity))
