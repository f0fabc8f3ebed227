"""cpuvoxelraycaster_b200 — B200-native voxel ray-traversal engine (Python mirror of the C ABI).

The product is libvrt.so (hand-written sm_100a CUDA behind include/vrt.h).  This package is the
host-side mirror of the reference's interface for the hot path — same names and argument meaning as
the reference's C++ classes (LSVO, Grid3D, MipmapGrid3D, SVO, RayCaster, Camera) — over ctypes.
It never computes on the CPU: without the built extension and a CUDA device every traversal call
raises.  (Scene construction on the host — terrain heights, octree flattening — is host code in the
reference too, src/main.cpp:59-86, and is exposed as `host_*`.)
"""
from . import capi
from .capi import HIT, LNODE, VrtError
from .engine import (Camera, CameraController, FlyController, ReplayElements, Context, Grid3D, LSVO, MipmapGrid3D, RayCaster, SVO, Volumetric, host_build_lsvo_from_voxels,
                     host_build_terrain_lsvo, host_terrain_heights)

__all__ = ["capi", "HIT", "LNODE", "VrtError", "Context", "Volumetric", "LSVO", "Grid3D", "MipmapGrid3D", "SVO",
           "RayCaster", "Camera", "CameraController", "FlyController", "ReplayElements", "host_terrain_heights", "host_build_terrain_lsvo", "host_build_lsvo_from_voxels"]
