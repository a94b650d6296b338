#include "cones_mat.cuh"
void hyp_mat_alloc_group(hyp_ctx*, ConeGroup&) { throw HypError{"matrix cones: not built yet"}; }
void hyp_mat_update_state(hyp_ctx*, ConeGroup&) { throw HypError{"matrix cones: not built yet"}; }
void hyp_mat_prod(hyp_ctx*, ConeGroup&, double*, const double*, int64_t, int64_t, int64_t, int, int64_t) {
    throw HypError{"matrix cones: not built yet"};
}
void hyp_mat_dder3(hyp_ctx*, ConeGroup&, double*, const double*) { throw HypError{"matrix cones: not built yet"}; }
