// spr_join.h -- the pair-join scorer: exact inlier count of EVERY lattice hypothesis from the
// (query landmark, reference landmark) pairs that can match at all.
//
// MatchMaps (place_recognition.cpp:98-387) counts, for every hypothesis (yaw, x, y), the query
// landmarks that have a reference landmark of the same label within match_threshold_ (and with
// matching dimensions).  Seen from a PAIR (q, r) under one yaw: the pair is an inlier exactly for the
// lattice translations within match_threshold_ of  r - R(yaw) q  -- a handful of lattice points.  The
// pair-join scorer enumerates those pairs through a coarse uniform grid over the reference map and
// adds each pair's hits to per-hypothesis counters in shared memory: work proportional to the sum of
// all inlier counts instead of (hypotheses x query landmarks).  The decision arithmetic is the
// reference's (spr_core.h: fp64, every operation rounded separately), so every counter ends at the
// reference's count; a query landmark with several matching reference landmarks is counted once, by
// the match with the lowest reference index (the reference's "first match, then break",
// PR.cpp:341-354).
//
//   block      : a rectangle of one ring's lattice (nx x ny samples) whose counters one CTA holds in
//                shared memory: two arrays of u16 counters, two per 32-bit word; word n of a row of array ay
//                holds samples 2n - ay and 2n - ay + 1, so that the two neighbouring samples a pair can hit
//                in a row always share ONE word of one array: one shared-memory atomic per pair and row.
//   quad       : 32 query landmarks (four query groups) handled by one warp: the candidate reference
//                landmarks of all 32 go through one work list so that the lanes stay busy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "spr_join_types.h"
#include "spr_types.h"

cudaError_t spr_launch_join_rotate(const SprJoinView &V, double *qrot, SprJoinBox *gbox, cudaStream_t st);
cudaError_t spr_launch_join_score(const SprJoinView &V, const SprJoinLaunch &K, int sm_count, cudaStream_t st);
// explicit hypothesis list (n x 4: c, s, x, y) against the prepared maps: inlier count of each, best (max count, lowest index)
cudaError_t spr_launch_join_score_list(const SprJoinView &V, const double *hyps4, long long n, int32_t *counts_out,
                                       unsigned long long *best_key, int sm_count, cudaStream_t st);
