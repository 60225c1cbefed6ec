"""Known-answer vectors HELD BY THE REFERENCE'S OWN TESTS for the CLIPPER half (restated here as
data; CSO = backend/sloam/clipper_semantic_object):
  * CSO/test/affinity_test.cpp:33-107   model / data clouds, the 12 x 12 affinity matrix `Mtrue`
  * CSO/test/clipper_test.cpp:15-68     the selected clique: 3 associations with A(i,0) == A(i,1)
  * CSO/test/dsd_test.cpp:15-80         20-node weighted graph, densest subgraph {3, 5, 12, 14, 15}
Shared by the CPU suite (oracle) and the GPU suite (product)."""
import numpy as np


def kat_clouds():
    """affinity_test.cpp:33-49: 4 model points, data = T_MD^-1 * model with the last point dropped."""
    model = np.array([[0.0, 0, 0], [2, 0, 0], [0, 3, 0], [2, 2, 0]]).T          # 3 x 4, points are columns
    a = np.pi / 8
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    t = np.array([5.0, 3.0, 0.0])
    data = R.T @ (model - t[:, None])                                             # T_MD.inverse() * model
    return model, data[:, :3]


MTRUE = np.array([
    [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0],
    [0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0],
    [0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0],
    [0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0],
    [1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 0, 0],
    [0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0],
    [0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0],
    [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0],
    [0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0],
    [0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0],
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1]], np.float64)                          # affinity_test.cpp:93-105

DSD_NODES = [3, 5, 12, 14, 15]                                                   # dsd_test.cpp:16
DSD_S = [0, 1, 3, 5, 7, 12, 14, 15, 19]                                          # dsd_test.cpp:72
_DSD_EDGES = {  # upper triangle of dsd_test.cpp:18-37 (zero entries omitted)
    (0, 18): 0.2964, (1, 13): 0.0138, (2, 11): 0.0016, (2, 18): 0.0747, (3, 5): 0.0555, (3, 6): 0.2547, (3, 13): 0.0102,
    (3, 15): 0.7715, (4, 5): 0.0063, (4, 7): 0.3846, (4, 9): 0.0003, (4, 10): 0.0014, (4, 15): 0.0063, (5, 12): 0.9927,
    (5, 15): 0.9722, (6, 8): 0.0023, (6, 11): 0.8775, (7, 8): 0.0001, (8, 9): 0.7914, (8, 13): 0.0617, (8, 16): 0.9938,
    (8, 19): 0.0007, (9, 12): 0.0001, (9, 13): 0.0091, (9, 15): 0.2503, (9, 16): 0.0222, (9, 17): 0.0549, (10, 19): 0.0008,
    (11, 18): 0.7007, (12, 14): 0.9978, (13, 17): 0.0003, (14, 15): 0.0012, (14, 19): 0.0074, (15, 16): 0.0026,
    (15, 17): 0.0217, (17, 18): 0.0007}


def dsd_matrix():
    M = np.eye(20)
    for (i, j), w in _DSD_EDGES.items():
        M[i, j] = M[j, i] = w
    return M
