// C ABI, part 2: Hyrax polynomial commitment (placeholder until the MSM kernels land)
#pragma once
#include "ctx.hpp"
extern "C" {
#define ZK_NOT_YET(name) { zk::g_last_error = name ": not implemented yet"; return -1; }
int zk_poly_bind_input(zk_ctx *, const uint64_t *, uint32_t) ZK_NOT_YET("zk_poly_bind_input")
int zk_poly_create(zk_ctx *, const uint64_t *, uint64_t, const uint64_t *, uint32_t) ZK_NOT_YET("zk_poly_create")
int zk_poly_commit(zk_ctx *, uint64_t *, uint32_t) ZK_NOT_YET("zk_poly_commit")
int zk_poly_evaluate(zk_ctx *, const uint64_t *, uint32_t, uint64_t *) ZK_NOT_YET("zk_poly_evaluate")
int zk_poly_init_bullet_prove(zk_ctx *, const uint64_t *, uint32_t, const uint64_t *, uint32_t) ZK_NOT_YET("zk_poly_init_bullet_prove")
int zk_poly_bullet_prove(zk_ctx *, uint64_t *, uint64_t *, uint64_t *, uint64_t *) ZK_NOT_YET("zk_poly_bullet_prove")
int zk_poly_bullet_update(zk_ctx *, const uint64_t *) ZK_NOT_YET("zk_poly_bullet_update")
int zk_poly_bullet_open(zk_ctx *, uint64_t *) ZK_NOT_YET("zk_poly_bullet_open")
}
