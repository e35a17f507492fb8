"""GPU: MapPlane::UpdateCoefficientsAndPoints (src/MapPlane.cc:298-365) through pslam_map_plane_update_batch vs the CPU oracle (oracle/planepost.cc
map_plane_update; PCL absent: parity unpinned): identical voxel sets in identical order, centroids bit-exact (order-free fixed-point sums on both sides)."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200.synth_mapplane import make_map_plane

pytestmark = pytest.mark.gpu


def test_map_plane_update_matches_oracle():
    from planarslam_b200._lib import Context, PslamError
    from planarslam_b200.planes import UpdateMapPlanePoints
    ctx = Context(640, 480, 1)
    jobs = [make_map_plane(0), make_map_plane(1, n_obs=1), make_map_plane(2, n_obs=12, pts_per_obs=900, extent=3.0), make_map_plane(3, with_current=True),
            make_map_plane(4, n_obs=3, pts_per_obs=5, extent=0.2), [], make_map_plane(5, n_obs=2, pts_per_obs=2000, extent=1.0, noise=0.03)]
    got = UpdateMapPlanePoints(ctx, jobs)
    assert len(got) == len(jobs)
    total = 0
    for j, clouds in enumerate(jobs):
        want = oracle_lib.map_plane_update(clouds)
        assert got[j].shape == want.shape, (j, got[j].shape, want.shape)
        assert np.array_equal(got[j].view(np.uint32), want.view(np.uint32)), j
        total += len(want)
    assert total > 1500 and len(got[5]) == 0
    # one job alone gives the same points as inside the batch
    alone = UpdateMapPlanePoints(ctx, [jobs[2]])[0]
    assert np.array_equal(alone, got[2])
    # capacity: fewer output slots than occupied voxels; more occupied voxels (4201) than the table holds (4096)
    with pytest.raises(PslamError):
        UpdateMapPlanePoints(ctx, [jobs[2]], cap=16)
    with pytest.raises(PslamError):
        UpdateMapPlanePoints(ctx, [make_map_plane(2, n_obs=12, pts_per_obs=900, extent=5.0)])
