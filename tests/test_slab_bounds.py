"""Host logic of the z-slab boundaries (icpb_slab_bounds_from_work, C-ABI; no GPU needed): every slab owns at least one
layer, the boundaries span the grid, and the shares of the work histogram are as equal as whole layers allow."""
import numpy as np
import pytest


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8, 16])
def test_bounds_are_valid_and_balanced(world):
    import icpb200
    rng = np.random.default_rng(world)
    for trial in range(40):
        layers = int(rng.integers(world, 600))
        kind = trial % 4
        if kind == 0:
            work = rng.integers(0, 1000, layers)
        elif kind == 1:      # everything in a few layers (a wall perpendicular to z)
            work = np.zeros(layers, np.int64)
            work[rng.integers(0, layers, 3)] = 10 ** 6
        elif kind == 2:
            work = np.zeros(layers, np.int64)
        else:                # smooth ramp
            work = (np.arange(layers) ** 2).astype(np.int64)
        b = icpb200.slab_bounds_from_work(work.astype(np.uint64), world)
        assert b[0] == 0 and b[-1] == layers and len(b) == world + 1
        assert all(b[g + 1] > b[g] for g in range(world)), b
        if kind in (0, 3) and layers >= 8 * world:
            shares = np.array([work[b[g]:b[g + 1]].sum() for g in range(world)], dtype=np.float64)
            # no slab exceeds the ideal share by more than the heaviest single layer (boundaries are whole layers)
            assert shares.max() <= work.sum() / world + work.max() + 1, (shares, b)


def test_bounds_match_equal_split_on_flat_work():
    import icpb200
    from icpb200 import dist as D
    work = np.full(500, 7, np.uint64)
    for world in (2, 4, 5):
        b = icpb200.slab_bounds_from_work(work, world)
        assert b == [D.shard_range(500, g, world)[0] for g in range(world)] + [500]
    b = icpb200.slab_bounds_from_work(work, 8)
    assert set(np.diff(b)) <= {62, 63}
