// kinds.h — static per-kind tables for the 25 constraint kinds: how many residual rows a kind has,
// which of the record's ids[] each row's sparsity list names (reference order, duplicates kept), and
// in which order the Jacobian kernels emit partial derivatives for each row.
//
// Restates, as data, ezpz/src/constraints.rs `residual_dim` (:954-993), `nonzeroes` (:378-491) and
// the JacobianVar emission order of `jacobian_rows` (:1000-2293).  ids[] layout: include/ezpz_b200.h.
#pragma once
#include <cstdint>

#include "../../include/ezpz_b200.h"

namespace ezk {

struct KindInfo {
    uint8_t n_ids;       // how many of ids[] are meaningful
    uint8_t rows;        // residual_dim
    uint8_t nz_len[2];   // length of each row's `nonzeroes` list
    uint8_t nz[2][8];    // indices into ids[]
    uint8_t emit_len[2]; // partials emitted per row by jacobian_rows (when not degenerate)
    uint8_t emit[2][8];  // indices into ids[], in emission order (accumulation order for duplicates)
};

// clang-format off
static const KindInfo kKinds[EZPZ_K_COUNT] = {
    /* 0  LineTangentToCircle  */ {7, 1, {7, 0}, {{0,1,2,3,4,5,6}, {}},            {7, 0}, {{0,1,2,3,4,5,6}, {}}},
    /* 1  CircleTangentToCircle*/ {6, 1, {6, 0}, {{0,1,2,3,4,5}, {}},              {6, 0}, {{0,1,2,3,4,5}, {}}},
    /* 2  Distance             */ {4, 1, {4, 0}, {{0,1,2,3}, {}},                  {4, 0}, {{0,1,2,3}, {}}},
    /* 3  DistanceVar          */ {5, 1, {5, 0}, {{0,1,2,3,4}, {}},                {5, 0}, {{0,1,2,3,4}, {}}},
    /* 4  VerticalDistance     */ {4, 1, {2, 0}, {{1,3}, {}},                      {2, 0}, {{1,3}, {}}},
    /* 5  HorizontalDistance   */ {4, 1, {2, 0}, {{0,2}, {}},                      {2, 0}, {{0,2}, {}}},
    /* 6  Vertical             */ {4, 1, {2, 0}, {{0,2}, {}},                      {2, 0}, {{0,2}, {}}},
    /* 7  Horizontal           */ {4, 1, {2, 0}, {{1,3}, {}},                      {2, 0}, {{1,3}, {}}},
    /* 8  LinesAtAngle         */ {8, 1, {8, 0}, {{0,1,2,3,4,5,6,7}, {}},          {8, 0}, {{0,1,2,3,4,5,6,7}, {}}},
    /* 9  Fixed                */ {1, 1, {1, 0}, {{0}, {}},                        {1, 0}, {{0}, {}}},
    /* 10 ScalarEqual          */ {2, 1, {2, 0}, {{0,1}, {}},                      {2, 0}, {{0,1}, {}}},
    /* 11 PointsCoincident     */ {4, 2, {2, 2}, {{0,2}, {1,3}},                   {2, 2}, {{0,2}, {1,3}}},
    /* 12 CircleRadius         */ {3, 1, {1, 0}, {{2}, {}},                        {1, 0}, {{2}, {}}},
    /* 13 LinesEqualLength     */ {8, 1, {8, 0}, {{0,1,2,3,4,5,6,7}, {}},          {8, 0}, {{0,1,2,3,4,5,6,7}, {}}},
    /* 14 ArcRadius            */ {6, 2, {4, 4}, {{4,5,0,1}, {4,5,2,3}},           {4, 4}, {{4,5,0,1}, {4,5,2,3}}},
    /* 15 Arc                  */ {6, 1, {6, 0}, {{0,1,2,3,4,5}, {}},              {6, 0}, {{0,1,2,3,4,5}, {}}},
    /* 16 Midpoint             */ {6, 2, {3, 3}, {{0,2,4}, {1,3,5}},               {3, 3}, {{4,0,2}, {5,1,3}}},
    /* 17 PointLineDistance    */ {6, 1, {6, 0}, {{0,1,2,3,4,5}, {}},              {6, 0}, {{0,1,2,3,4,5}, {}}},
    /* 18 VerticalPointLineDist*/ {6, 1, {6, 0}, {{2,3,4,5,0,1}, {}},              {6, 0}, {{0,1,2,3,4,5}, {}}},
    /* 19 HorizontalPointLineD.*/ {6, 1, {6, 0}, {{2,3,4,5,0,1}, {}},              {6, 0}, {{0,1,2,3,4,5}, {}}},
    /* 20 Symmetric            */ {8, 2, {8, 8}, {{0,1,2,3,4,5,6,7}, {0,1,2,3,4,5,6,7}}, {8, 8}, {{0,1,2,3,4,5,6,7}, {0,1,2,3,4,5,6,7}}},
    /* 21 PointArcCoincident   */ {8, 2, {8, 8}, {{0,1,2,3,4,5,6,7}, {0,1,2,3,4,5,6,7}}, {8, 8}, {{4,5,0,1,2,3,6,7}, {4,5,0,1,2,3,6,7}}},
    /* 22 ArcLength            */ {6, 2, {6, 6}, {{0,1,2,3,4,5}, {0,1,2,3,4,5}},   {6, 6}, {{0,1,2,3,4,5}, {0,1,2,3,4,5}}},
    /* 23 ArcAngle             */ {6, 1, {8, 0}, {{4,5,0,1,4,5,2,3}, {}},          {8, 0}, {{4,5,0,1,4,5,2,3}, {}}},
    /* 24 PointsAtAngle        */ {6, 2, {6, 6}, {{0,1,2,3,4,5}, {0,1,2,3,4,5}},   {6, 6}, {{0,1,2,3,4,5}, {0,1,2,3,4,5}}},
};
// clang-format on

inline const KindInfo* kind_info(uint32_t kind) { return kind < EZPZ_K_COUNT ? &kKinds[kind] : nullptr; }

}  // namespace ezk
