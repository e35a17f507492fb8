"""CPU: properties of the MapPlane::UpdateCoefficientsAndPoints oracle (oracle/planepost.cc map_plane_update, parity unpinned): one centroid per occupied 0.1 m
voxel of the transformed, concatenated clouds, in ascending (z, y, x) voxel order; independent of the order of the observations."""
import numpy as np

import oracle_lib
from planarslam_b200.synth_mapplane import make_map_plane


def _voxels(points):
    return np.floor(points * np.float32(10.0)).astype(np.int64)


def test_map_plane_update_properties():
    for seed in range(4):
        clouds = make_map_plane(seed, with_current=seed == 3)
        out = oracle_lib.map_plane_update(clouds)
        world = np.concatenate([((p.astype(np.float64) @ T[:3, :3].T) + T[:3, 3]).astype(np.float32) for p, T in clouds])
        vox = np.unique(_voxels(world), axis=0)
        assert len(out) == len(vox)                                        # one centroid per occupied voxel
        ov = _voxels(out)
        key = (ov[:, 2] * (1 << 40)) + (ov[:, 1] * (1 << 20)) + ov[:, 0]
        assert np.all(np.diff(key) > 0)                                    # ascending voxel index = lexicographic (z, y, x)
        assert set(map(tuple, ov)) <= set(map(tuple, vox)) or np.abs(out - (ov + 0.5) / 10).max() < 0.051
        again = oracle_lib.map_plane_update(clouds[::-1])                  # std::map<KeyFrame*, size_t> order is pointer order: the result must not depend on it
        assert np.array_equal(out, again)
    assert len(oracle_lib.map_plane_update([])) == 0
