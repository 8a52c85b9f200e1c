// camera_host.cpp — once-per-frame host setup of the pinhole PerspectiveCamera.
//
// Follows src/cameras/perspective.rs:34-82 and src/core/transform.rs (perspective :555-566, scale :426-443,
// translate :408-425, look_at :510-541, Matrix4x4::inverse :46-113, Transform product :609-617) with the Appendix-A
// fixes (D4: Gauss-Jordan starts from the matrix; D5: (AB)^-1 = B^-1 A^-1; scale() builds a diagonal).  f32 throughout,
// no FMA contraction (-ffp-contract=off), so the matrices equal the reference arithmetic's bit for bit.
#include <cmath>
#include <cstring>
#include <utility>

#include "kernels.hpp"

namespace pb2 {
namespace {

struct Xf {
    mat4 m, inv;
};

mat4 ident() {
    mat4 r;
    std::memset(&r, 0, sizeof r);
    r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.0f;
    return r;
}

mat4 matmul(const mat4& a, const mat4& b) {
    mat4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float acc = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j];
            acc = acc + a.m[i][2] * b.m[2][j];
            r.m[i][j] = acc + a.m[i][3] * b.m[3][j];
        }
    return r;
}

// Gauss-Jordan elimination with full pivoting (pbrt's Inverse()).
mat4 invert(const mat4& src) {
    float w[4][4];
    std::memcpy(w, src.m, sizeof w);
    int col_of[4], row_of[4], used[4] = {0, 0, 0, 0};
    for (int step = 0; step < 4; ++step) {
        int pr = 0, pc = 0;
        float best = 0.0f;
        for (int r = 0; r < 4; ++r) {
            if (used[r] == 1) continue;
            for (int c = 0; c < 4; ++c)
                if (used[c] == 0 && std::fabs(w[r][c]) >= best) { best = std::fabs(w[r][c]); pr = r; pc = c; }
        }
        used[pc] += 1;
        if (pr != pc)
            for (int k = 0; k < 4; ++k) std::swap(w[pr][k], w[pc][k]);
        row_of[step] = pr;
        col_of[step] = pc;
        const float pivinv = 1.0f / w[pc][pc];
        w[pc][pc] = 1.0f;
        for (int k = 0; k < 4; ++k) w[pc][k] *= pivinv;
        for (int r = 0; r < 4; ++r) {
            if (r == pc) continue;
            const float f = w[r][pc];
            w[r][pc] = 0.0f;
            for (int k = 0; k < 4; ++k) w[r][k] -= w[pc][k] * f;
        }
    }
    for (int step = 3; step >= 0; --step)
        if (row_of[step] != col_of[step])
            for (int k = 0; k < 4; ++k) std::swap(w[k][row_of[step]], w[k][col_of[step]]);
    mat4 out;
    std::memcpy(out.m, w, sizeof w);
    return out;
}

Xf compose(const Xf& a, const Xf& b) { return Xf{matmul(a.m, b.m), matmul(b.inv, a.inv)}; }
Xf flipped(const Xf& a) { return Xf{a.inv, a.m}; }

Xf scaling(float x, float y, float z) {
    Xf t{ident(), ident()};
    t.m.m[0][0] = x; t.m.m[1][1] = y; t.m.m[2][2] = z;
    t.inv.m[0][0] = 1.0f / x; t.inv.m[1][1] = 1.0f / y; t.inv.m[2][2] = 1.0f / z;
    return t;
}

Xf translation(float x, float y, float z) {
    Xf t{ident(), ident()};
    t.m.m[0][3] = x; t.m.m[1][3] = y; t.m.m[2][3] = z;
    t.inv.m[0][3] = -x; t.inv.m[1][3] = -y; t.inv.m[2][3] = -z;
    return t;
}

Xf perspective(float fov_deg, float n, float f) {
    mat4 p = ident();
    p.m[2][2] = f / (f - n);
    p.m[2][3] = -f * n / (f - n);
    p.m[3][2] = 1.0f;
    p.m[3][3] = 0.0f;
    const float inv_tan = 1.0f / std::tan((PB2_PI / 180.0f * fov_deg) / 2.0f);
    return compose(scaling(inv_tan, inv_tan, 1.0f), Xf{p, invert(p)});
}

}  // namespace

mat4 invert_mat4(const mat4& m) { return invert(m); }

// cam9 = pos, look, up
void camera_setup(const float pos[3], const float look[3], const float up[3], float fov, int res_x, int res_y, CameraView* out) {
    out->res_x = res_x;
    out->res_y = res_y;
    // pbrt's default screen window: the shorter image axis spans [-1, 1]
    const float aspect = (float)res_x / (float)res_y;
    float x0, x1, y0, y1;
    if (aspect > 1.0f) { x0 = -aspect; x1 = aspect; y0 = -1.0f; y1 = 1.0f; }
    else { x0 = -1.0f; x1 = 1.0f; y0 = -1.0f / aspect; y1 = 1.0f / aspect; }
    const Xf camera_to_screen = perspective(fov, 1e-2f, 1000.0f);
    const Xf screen_to_raster = compose(compose(scaling((float)res_x, (float)res_y, 1.0f), scaling(1.0f / (x1 - x0), 1.0f / (y0 - y1), 1.0f)),
                                        translation(-x0, -y1, 0.0f));
    const Xf raster_to_camera = compose(flipped(camera_to_screen), flipped(screen_to_raster));
    out->raster_to_camera = raster_to_camera.m;

    // Transform::look_at: build camera_to_world, the transform's m is its inverse
    const vec3 eye = mk(pos[0], pos[1], pos[2]);
    const vec3 dir = unit(mk(look[0], look[1], look[2]) - eye);
    const vec3 right = unit(cross3(unit(mk(up[0], up[1], up[2])), dir));
    const vec3 new_up = cross3(dir, right);
    mat4 c2w = ident();
    c2w.m[0][0] = right.x; c2w.m[1][0] = right.y; c2w.m[2][0] = right.z;
    c2w.m[0][1] = new_up.x; c2w.m[1][1] = new_up.y; c2w.m[2][1] = new_up.z;
    c2w.m[0][2] = dir.x; c2w.m[1][2] = dir.y; c2w.m[2][2] = dir.z;
    c2w.m[0][3] = eye.x; c2w.m[1][3] = eye.y; c2w.m[2][3] = eye.z;
    // the camera's CameraToWorld is Inverse(LookAt) = {m: c2w, inv: invert(c2w)}
    out->camera_to_world = c2w;
}

}  // namespace pb2
