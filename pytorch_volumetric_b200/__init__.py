"""pytorch_volumetric_b200 -- B200-native (sm_100a) drop-in for pytorch_volumetric's batched SDF query path.

Same public names as the reference's `pytorch_volumetric/__init__.py` for that path.
"""
from .chamfer import batch_chamfer_dist, PlausibleDiversity, pairwise_distance, pairwise_distance_chamfer
from .sdf import sample_mesh_points, ObjectFrameSDF, MeshSDF, CachedSDF, ComposedSDF, SDFQuery, \
    ObjectFactory, MeshObjectFactory, OutOfBoundsStrategy, SphereSDF
from .voxel import Voxels, VoxelGrid, VoxelSet, ExpandingVoxelGrid, get_divisible_range_by_resolution, \
    get_coordinates_and_points_in_grid, voxel_down_sample
from .model_to_sdf import RobotSDF, cache_link_sdf_factory, aabb_to_ordered_end_points
from .transforms import Transform3d, Translate
from .kinematics import build_serial_chain_from_urdf, SerialChain
from . import distributed

__all__ = [
    "batch_chamfer_dist", "PlausibleDiversity", "pairwise_distance", "pairwise_distance_chamfer",
    "sample_mesh_points", "ObjectFrameSDF", "MeshSDF", "CachedSDF", "ComposedSDF", "SDFQuery", "ObjectFactory",
    "MeshObjectFactory", "OutOfBoundsStrategy", "SphereSDF", "Voxels", "VoxelGrid", "VoxelSet", "ExpandingVoxelGrid",
    "voxel_down_sample", "get_divisible_range_by_resolution",
    "get_coordinates_and_points_in_grid", "RobotSDF", "cache_link_sdf_factory", "aabb_to_ordered_end_points",
    "Transform3d", "Translate", "build_serial_chain_from_urdf", "SerialChain", "distributed",
]
