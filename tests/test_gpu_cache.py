"""GPU suite (-m gpu): the device-resident map cache keyed by robot id (SURVEY.md section 8f-3;
databaseManager::robotMapDict_, databaseManager.h:99-102) and the batch entry built on it.  A cached
search must return exactly what findTransformation returns on the same maps handed over by value,
however the pairs interleave."""
import numpy as np
import pytest

import spr_helpers as H
from slide_slam_b200 import capi, synth
from slide_slam_b200.place_recognition import PlaceRecognition

pytestmark = pytest.mark.gpu

ROS = {"search_xy_step_size": 0.5, "search_yaw_step_size_degrees": 5.0, "match_threshold_position": 0.5,
       "match_threshold_dimension": 1.0, "ignore_dimension": 0, "min_num_inliers": 15}


def same(a, b):
    fa, xa, ta, ia, ra, qa = a
    fb, xb, tb, ib, rb, qb = b
    assert fa == fb and ia.best_num_inliers == ib.best_num_inliers and ia.match.best_hyp_index == ib.match.best_hyp_index
    assert ra.tolist() == rb.tolist() and qa.tolist() == qb.tolist()
    assert list(ia.R_t) == list(ib.R_t) and xa.tolist() == xb.tolist() and ta.tolist() == tb.tolist()
    assert ia.half_x == ib.half_x and list(ia.centroid_ref) == list(ib.centroid_ref) and list(ia.centroid_qry) == list(ib.centroid_qry)


def test_cached_search_equals_search_by_value():
    maps = synth.config_robots(4, 600)
    plain, pr = PlaceRecognition(ROS), PlaceRecognition(ROS)
    for i, m in enumerate(maps):
        pr.cache_put(10 + i, 1, m)
    assert pr.cache_size() == 4
    order = [(0, 1), (2, 1), (0, 3), (2, 3), (0, 1), (1, 0), (2, 0)]     # references interleave: a single slot would thrash
    seen_ref = set()
    for r, q in order:
        got = pr.findTransformationCached(10 + r, 10 + q)
        same(got, plain.findTransformation(maps[r], maps[q]))
        assert bool(got[3].match.reuse & 2) == (r in seen_ref)         # the index of a reference map is built once
        seen_ref.add(r)
    # a new version of map 0 rebuilds its index; the same version does not
    moved = maps[0].copy(); moved[:, 1] += 3.0
    pr.cache_put(10, 2, moved)
    got = pr.findTransformationCached(10, 11)
    assert not (got[3].match.reuse & 2)
    same(got, plain.findTransformation(moved, maps[1]))
    pr.cache_put(10, 2, moved)
    assert pr.findTransformationCached(10, 11)[3].match.reuse & 2
    # drop / unknown ids
    assert pr.cache_drop(13) and not pr.cache_drop(13) and pr.cache_size() == 3
    with pytest.raises(capi.SlidePrError):
        pr.findTransformationCached(10, 13)
    # the by-value entry points keep working next to the cache (anonymous slot)
    same(pr.findTransformation(maps[2], maps[3]), plain.findTransformation(maps[2], maps[3]))
    same(pr.findTransformationCached(11, 12), plain.findTransformation(maps[1], maps[2]))
    pr.close(); plain.close()


def test_batch_equals_individual_calls_and_builds_each_index_once():
    maps = synth.config_robots(5, 500)
    pairs = [(i, j) for i in range(5) for j in range(i + 1, 5)]
    pr, plain = PlaceRecognition(ROS), PlaceRecognition(ROS)
    outs = pr.findTransformationBatch(maps, pairs)
    built = 0
    for (r, q), o in zip(pairs, outs):
        f, x, t, info, ri, qi = plain.findTransformation(maps[r], maps[q])
        assert bool(o.found) == f and o.best_num_inliers == info.best_num_inliers and o.match.best_hyp_index == info.match.best_hyp_index
        assert list(o.xyz_yaw) == x.tolist() and list(o.transform) == t.ravel().tolist()
        built += int(not (o.match.reuse & 2))
    assert built == 4                     # maps 0..3 serve as references: one index each, 10 pairs
    assert pr.cache_size() == 0           # the batch's private slots are gone
    pr.close(); plain.close()


def test_cache_eviction_keeps_results_right():
    maps = synth.config_robots(3, 300)
    pr, plain = PlaceRecognition(ROS), PlaceRecognition(ROS)
    for k in range(70):                   # more slots than the cache keeps (64): the oldest are dropped
        pr.cache_put(1000 + k, 1, maps[k % 3])
    assert pr.cache_size() == 64
    with pytest.raises(capi.SlidePrError):
        pr.findTransformationCached(1000, 1001)
    same(pr.findTransformationCached(1068, 1069), plain.findTransformation(maps[68 % 3], maps[69 % 3]))
    pr.close(); plain.close()
