// q3_plan.h — the TPC-H-Q3-shaped delta-join plan as closure descriptors: the
// workload definition shared by the GPU harness (harness.cu) and the CPU oracle
// (oracle/dataflow.cc), so both sides execute the same rendered plan.
//
// Plan source: test/sqllogictest/tpch_create_materialized_view.slt:320-369
//   %0:customer » %1:orders[#1{o_custkey}]KAif » %2:lineitem[#0{l_orderkey}]KAif
//   %1:orders   » %0:customer[#0{c_custkey}]KAef » %2:lineitem[#0{l_orderkey}]KAif
//   %2:lineitem » %1:orders[#0{o_orderkey}]KAif » %0:customer[#0{c_custkey}]KAef
// Shape: DeltaPathPlan / DeltaStagePlan, src/compute-types/src/plan/join/delta_join.rs:46-80.
// Column packing of the value words: gen.h.
#pragma once
#include <string.h>

#include "gen.h"

static inline mzgpu_field mzg_F(uint8_t src, uint8_t shift, uint8_t bits, uint8_t dst) {
  mzgpu_field f;
  f.src = src;
  f.shift = shift;
  f.bits = bits;
  f.dst_shift = dst;
  return f;
}
static inline mzgpu_filter mzg_FL(uint8_t src, uint8_t shift, uint8_t bits, uint32_t op, uint64_t rhs) {
  mzgpu_filter f;
  f.field = mzg_F(src, shift, bits, 0);
  f.op = op;
  f.rhs = rhs;
  return f;
}

// The Q3 plan as closure descriptors; shared verbatim with the GPU harness
// through mzo_q3_plan() so both sides run the same plan.
typedef struct mzg_q3_plan {
  // [path] initial closure, [path][stage] closure, cmp mode, lookup arrangement
  mzgpu_closure initial[3];
  mzgpu_closure stage[3][2];
  int32_t cmp[3][2];
  int32_t lookup[3][2];  // 0 = customer[custkey], 1 = orders[orderkey], 2 = orders[custkey], 3 = lineitem[orderkey]
  int32_t source[3];     // arrangement whose batches feed the path
} mzg_q3_plan;

static inline void mzg_q3_plan_init(mzg_q3_plan* p) {
  memset(p, 0, sizeof(*p));
  const uint64_t CUT = MZG_Q3_DATE_CUTOFF;
  // ---- path 0: customer » orders[custkey] » lineitem[orderkey]
  p->source[0] = 0;
  {
    mzgpu_closure& c = p->initial[0];  // filter c_mktsegment = BUILDING; keep key
    c.n_filters = 1;
    c.filters[0] = mzg_FL(MZGPU_SRC_VAL1, 0, 3, MZGPU_CMP_EQ, MZG_Q3_SEGMENT);
    c.n_key_fields = 1;
    c.key_fields[0] = mzg_F(MZGPU_SRC_KEY, 0, 64, 0);
  }
  {
    mzgpu_closure& c = p->stage[0][0];  // lookup orders by custkey, o_orderdate < cutoff
    c.n_filters = 1;
    c.filters[0] = mzg_FL(MZGPU_SRC_VAL2, 32, 12, MZGPU_CMP_LT, CUT);
    c.n_key_fields = 1;
    c.key_fields[0] = mzg_F(MZGPU_SRC_VAL2, 0, 32, 0);   // next key: orderkey
    c.n_val_fields = 1;
    c.val_fields[0] = mzg_F(MZGPU_SRC_VAL2, 32, 13, 0);  // orderdate | shippriority << 12
    p->cmp[0][0] = MZGPU_HALFJOIN_LE;
    p->lookup[0][0] = 2;
  }
  {
    mzgpu_closure& c = p->stage[0][1];  // lookup lineitem by orderkey, l_shipdate > cutoff
    c.n_filters = 1;
    c.filters[0] = mzg_FL(MZGPU_SRC_VAL2, 24, 12, MZGPU_CMP_GT, CUT);
    c.n_key_fields = 2;                 // group key: orderkey | (orderdate, shippriority) << 32
    c.key_fields[0] = mzg_F(MZGPU_SRC_KEY, 0, 32, 0);
    c.key_fields[1] = mzg_F(MZGPU_SRC_VAL1, 0, 13, 32);
    c.expr_kind = MZGPU_EXPR_MUL_CONST_MINUS;  // l_extendedprice * (100 - l_discount)
    c.expr_a = mzg_F(MZGPU_SRC_VAL2, 3, 17, 0);
    c.expr_b = mzg_F(MZGPU_SRC_VAL2, 20, 4, 0);
    c.expr_c = 100;
    p->cmp[0][1] = MZGPU_HALFJOIN_LE;
    p->lookup[0][1] = 3;
  }
  // ---- path 1: orders » customer[custkey] » lineitem[orderkey]
  p->source[1] = 1;
  {
    mzgpu_closure& c = p->initial[1];  // o_orderdate < cutoff; key := custkey
    c.n_filters = 1;
    c.filters[0] = mzg_FL(MZGPU_SRC_VAL1, 24, 12, MZGPU_CMP_LT, CUT);
    c.n_key_fields = 1;
    c.key_fields[0] = mzg_F(MZGPU_SRC_VAL1, 0, 24, 0);
    c.n_val_fields = 2;                // orderkey | (orderdate, shippriority) << 32
    c.val_fields[0] = mzg_F(MZGPU_SRC_KEY, 0, 32, 0);
    c.val_fields[1] = mzg_F(MZGPU_SRC_VAL1, 24, 13, 32);
  }
  {
    mzgpu_closure& c = p->stage[1][0];  // lookup customer, c_mktsegment = BUILDING
    c.n_filters = 1;
    c.filters[0] = mzg_FL(MZGPU_SRC_VAL2, 0, 3, MZGPU_CMP_EQ, MZG_Q3_SEGMENT);
    c.n_key_fields = 1;
    c.key_fields[0] = mzg_F(MZGPU_SRC_VAL1, 0, 32, 0);   // next key: orderkey
    c.n_val_fields = 1;
    c.val_fields[0] = mzg_F(MZGPU_SRC_VAL1, 32, 13, 0);
    p->cmp[1][0] = MZGPU_HALFJOIN_LT;  // orders(1) > customer(0)
    p->lookup[1][0] = 0;
  }
  p->stage[1][1] = p->stage[0][1];
  p->cmp[1][1] = MZGPU_HALFJOIN_LE;
  p->lookup[1][1] = 3;
  // ---- path 2: lineitem » orders[orderkey] » customer[custkey]
  p->source[2] = 3;
  {
    mzgpu_closure& c = p->initial[2];  // l_shipdate > cutoff; keep (extendedprice, discount)
    c.n_filters = 1;
    c.filters[0] = mzg_FL(MZGPU_SRC_VAL1, 24, 12, MZGPU_CMP_GT, CUT);
    c.n_key_fields = 1;
    c.key_fields[0] = mzg_F(MZGPU_SRC_KEY, 0, 64, 0);
    c.n_val_fields = 1;
    c.val_fields[0] = mzg_F(MZGPU_SRC_VAL1, 3, 21, 0);   // extprice[0:17] | discount[17:21]
  }
  {
    mzgpu_closure& c = p->stage[2][0];  // lookup orders by orderkey, o_orderdate < cutoff
    c.n_filters = 1;
    c.filters[0] = mzg_FL(MZGPU_SRC_VAL2, 24, 12, MZGPU_CMP_LT, CUT);
    c.n_key_fields = 1;
    c.key_fields[0] = mzg_F(MZGPU_SRC_VAL2, 0, 24, 0);   // next key: custkey
    c.n_val_fields = 3;
    c.val_fields[0] = mzg_F(MZGPU_SRC_VAL1, 0, 21, 0);   // extprice | discount
    c.val_fields[1] = mzg_F(MZGPU_SRC_VAL2, 24, 13, 21); // orderdate | shippriority
    c.val_fields[2] = mzg_F(MZGPU_SRC_KEY, 0, 30, 34);   // orderkey
    p->cmp[2][0] = MZGPU_HALFJOIN_LT;  // lineitem(2) > orders(1)
    p->lookup[2][0] = 1;
  }
  {
    mzgpu_closure& c = p->stage[2][1];  // lookup customer, c_mktsegment = BUILDING
    c.n_filters = 1;
    c.filters[0] = mzg_FL(MZGPU_SRC_VAL2, 0, 3, MZGPU_CMP_EQ, MZG_Q3_SEGMENT);
    c.n_key_fields = 2;
    c.key_fields[0] = mzg_F(MZGPU_SRC_VAL1, 34, 30, 0);
    c.key_fields[1] = mzg_F(MZGPU_SRC_VAL1, 21, 13, 32);
    c.expr_kind = MZGPU_EXPR_MUL_CONST_MINUS;
    c.expr_a = mzg_F(MZGPU_SRC_VAL1, 0, 17, 0);
    c.expr_b = mzg_F(MZGPU_SRC_VAL1, 17, 4, 0);
    c.expr_c = 100;
    p->cmp[2][1] = MZGPU_HALFJOIN_LT;  // lineitem(2) > customer(0)
    p->lookup[2][1] = 0;
  }
}
