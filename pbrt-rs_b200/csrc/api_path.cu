// api_path.cu — Film, wavefront PathIntegrator and NCCL entry points of include/pbrt_b200.h.
#include "api_internal.hpp"

namespace pb2 {
struct Wavefront {};
void wavefront_destroy(Wavefront* wf) { delete wf; }
int upload_shading_tables(pb2_scene*) { return PB2_OK; }
}  // namespace pb2

using namespace pb2;

extern "C" {
#define PB2_TODO(name) return set_error(PB2_ERR_STATE, name " is not implemented yet")
int pb2_film_create(const pb2_film_desc*, pb2_film**) { PB2_TODO("pb2_film_create"); }
int pb2_film_destroy(pb2_film*) { return PB2_OK; }
int pb2_film_clear(pb2_film*) { PB2_TODO("pb2_film_clear"); }
int pb2_film_add_samples(pb2_film*, const float*, const float*, const float*, uint64_t) { PB2_TODO("pb2_film_add_samples"); }
int pb2_film_read_xyzw(pb2_film*, float*) { PB2_TODO("pb2_film_read_xyzw"); }
int pb2_film_resolve_rgb(pb2_film*, float, float*) { PB2_TODO("pb2_film_resolve_rgb"); }
int pb2_film_device_ptr(pb2_film*, void**, uint64_t*) { PB2_TODO("pb2_film_device_ptr"); }
int pb2_render_path(pb2_scene*, const pb2_camera*, const pb2_path_desc*, pb2_film*, void*) { PB2_TODO("pb2_render_path"); }
int pb2_path_li(pb2_scene*, const pb2_camera*, const pb2_path_desc*, const uint32_t*, const uint32_t*, uint64_t, float*, float*) { PB2_TODO("pb2_path_li"); }
int pb2_render_counters(pb2_scene*, uint64_t*) { PB2_TODO("pb2_render_counters"); }
int pb2_nccl_unique_id(char*) { PB2_TODO("pb2_nccl_unique_id"); }
int pb2_nccl_init(const char*, int, int) { PB2_TODO("pb2_nccl_init"); }
int pb2_nccl_shutdown(void) { return PB2_OK; }
int pb2_film_reduce(pb2_film*, int, void*) { PB2_TODO("pb2_film_reduce"); }
}
