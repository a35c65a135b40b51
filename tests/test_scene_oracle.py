"""CPU: the scene-mode oracle (oracle/ref_scene.py, nocs/zero_shot.ipynb cells 6 and 9) on constructed inputs."""
import numpy as np

from oracle import ref_scene as rs


def test_proposals_find_separated_peaks_in_order_and_stop_on_contrast():
    g = rs.blob_grid((48, 40, 44), [(12, 10, 11), (34, 28, 30), (24, 20, 5)], [400.0, 350.0, 60.0])
    props = rs.proposals(g.copy(), thresh=50, margin=10)
    # the weak third peak (50 < contrast 60 < 0.7 * 399) is still appended, THEN the loop stops -- cell 9's order of tests
    assert [tuple(p[0]) for p in props] == [(12, 10, 11), (34, 28, 30), (24, 20, 5)]
    assert props[0][2] > props[1][2] > props[2][2] > 50
    g4 = rs.blob_grid((48, 40, 44), [(12, 10, 11), (34, 28, 30), (24, 20, 5), (40, 5, 38)], [400.0, 350.0, 60.0, 55.0])
    assert len(rs.proposals(g4, thresh=50, margin=10)) == 3                    # nothing after the stop
    assert rs.proposals(rs.blob_grid((30, 30, 30), [(15, 15, 15)], [20.0]), thresh=50) == []


def test_pair_filter_drops_only_coplanar_parallel_pairs():
    pc = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1], [1, 1, 0.3]], np.float32)
    nrm = np.array([[0, 0, 1], [0, 0, 1], [0, 0, 1], [1, 0, 0]], np.float32)
    idx = np.array([[0, 1], [0, 2], [0, 3], [1, 3]])
    np.testing.assert_array_equal(rs.pair_filter(pc, nrm, idx), [False, True, True, True])


def test_instance_points_threshold():
    pairs = np.array([[0, 1]] * 7 + [[2, 3]] * 3)
    keep_pt, keep_pair = rs.instance_points(pairs, 5, min_contrib=6)
    np.testing.assert_array_equal(keep_pt, [True, True, False, False, False])
    assert keep_pair.sum() == 7
