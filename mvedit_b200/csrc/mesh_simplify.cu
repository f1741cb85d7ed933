// mesh_simplify.cu -- quadric-error edge-collapse decimation of the DMTet mesh, HOST code (no kernel in this file).
//
// Where the reference calls it: the last mesh_optim of a run with mesh_reduction < 1 hands the marching-tets mesh to open3d's
// TriangleMesh.simplify_quadric_decimation(round(n_faces * mesh_reduction), boundary_weight=0) ON THE CPU and then only fits the texture
// on the fixed, decimated mesh (/root/reference/lib/pipelines/mvedit_3d_pipeline.py:829-844; the runner asks for it when
// tet_resolution > 128, lib/apis/adapter3d.py:822).  open3d is a third-party dependency absent from this image, so this is the published
// algorithm (Garland & Heckbert 1997, as open3d 0.18 implements it: area-weighted plane quadrics, optimal placement by the 3x3 solve
// with end-point / mid-point candidates when singular, normal-flip rejection, lazy-deletion priority queue) -- "parity unpinned": the
// result is checked through properties (face count, manifoldness, orientation, distance to the input surface), not against open3d.
// One addition: the link condition (an interior edge collapses only if its end points share exactly two neighbours), which keeps a
// closed 2-manifold closed and manifold -- the textured mesh is rasterised and UV-mapped afterwards.
// It runs once per pipeline call, on ~10^5 faces, in well under a second: not a GPU job.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <queue>
#include <vector>

#include "../../include/mvedit_b200.h"

namespace {

struct Quadric {
    double q[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};      // a00 a01 a02 a03 a11 a12 a13 a22 a23 a33
    void add_plane(double a, double b, double c, double d, double w) {
        q[0] += w * a * a; q[1] += w * a * b; q[2] += w * a * c; q[3] += w * a * d;
        q[4] += w * b * b; q[5] += w * b * c; q[6] += w * b * d;
        q[7] += w * c * c; q[8] += w * c * d; q[9] += w * d * d;
    }
    void add(const Quadric& o) { for (int i = 0; i < 10; i++) q[i] += o.q[i]; }
    double eval(const double* v) const {
        const double x = v[0], y = v[1], z = v[2];
        return q[0] * x * x + 2 * q[1] * x * y + 2 * q[2] * x * z + 2 * q[3] * x + q[4] * y * y + 2 * q[5] * y * z + 2 * q[6] * y +
               q[7] * z * z + 2 * q[8] * z + q[9];
    }
    bool minimiser(double* v) const {                     // solve A v = -b (Cramer); false when A is (near) singular
        const double a = q[0], b = q[1], c = q[2], d = q[4], e = q[5], f = q[7];
        const double det = a * (d * f - e * e) - b * (b * f - e * c) + c * (b * e - d * c);
        const double scale = std::fabs(a) + std::fabs(d) + std::fabs(f);
        if (std::fabs(det) <= 1e-9 * scale * scale * scale || scale == 0) return false;
        const double r0 = -q[3], r1 = -q[6], r2 = -q[8];
        v[0] = (r0 * (d * f - e * e) - b * (r1 * f - e * r2) + c * (r1 * e - d * r2)) / det;
        v[1] = (a * (r1 * f - e * r2) - r0 * (b * f - e * c) + c * (b * r2 - r1 * c)) / det;
        v[2] = (a * (d * r2 - r1 * e) - b * (b * r2 - r1 * c) + r0 * (b * e - d * c)) / det;
        return true;
    }
};

struct Candidate {
    double cost;
    uint32_t v0, v1, s0, s1;                              // s*: version stamps of the end points when the candidate was made
    bool operator<(const Candidate& o) const { return cost > o.cost; }      // min-heap
};

struct Simplifier {
    std::vector<double> pos;                              // [V,3]
    std::vector<int32_t> tri;                             // [F,3]
    std::vector<char> face_alive, vert_alive;
    std::vector<uint32_t> stamp;
    std::vector<Quadric> Q;
    std::vector<std::vector<uint32_t>> vfaces;            // faces incident to a vertex (may hold dead faces: filtered on use)
    std::priority_queue<Candidate> heap;
    uint32_t n_faces = 0;

    void face_normal(uint32_t f, const double* moved, int32_t who, double* n) const {
        const double* p[3];
        for (int k = 0; k < 3; k++) p[k] = (tri[f * 3 + k] == who && moved) ? moved : &pos[(size_t)tri[f * 3 + k] * 3];
        const double ux = p[1][0] - p[0][0], uy = p[1][1] - p[0][1], uz = p[1][2] - p[0][2];
        const double vx = p[2][0] - p[0][0], vy = p[2][1] - p[0][1], vz = p[2][2] - p[0][2];
        n[0] = uy * vz - uz * vy; n[1] = uz * vx - ux * vz; n[2] = ux * vy - uy * vx;
    }

    void neighbours(uint32_t v, std::vector<uint32_t>& out) const {
        out.clear();
        for (uint32_t f : vfaces[v]) {
            if (!face_alive[f]) continue;
            for (int k = 0; k < 3; k++) {
                const uint32_t w = (uint32_t)tri[f * 3 + k];
                if (w != v) out.push_back(w);
            }
        }
        std::sort(out.begin(), out.end());
        out.erase(std::unique(out.begin(), out.end()), out.end());
    }

    double placement(uint32_t v0, uint32_t v1, double* best) const {
        Quadric q = Q[v0];
        q.add(Q[v1]);
        double cand[3];
        if (q.minimiser(cand)) {
            best[0] = cand[0]; best[1] = cand[1]; best[2] = cand[2];
            return std::max(q.eval(cand), 0.0);
        }
        const double* a = &pos[(size_t)v0 * 3];
        const double* b = &pos[(size_t)v1 * 3];
        const double mid[3] = {(a[0] + b[0]) * 0.5, (a[1] + b[1]) * 0.5, (a[2] + b[2]) * 0.5};
        const double* opts[3] = {mid, a, b};
        double cost = 1e300;
        for (const double* o : opts) {
            const double c = q.eval(o);
            if (c < cost) { cost = c; best[0] = o[0]; best[1] = o[1]; best[2] = o[2]; }
        }
        return std::max(cost, 0.0);
    }

    void push(uint32_t a, uint32_t b) {
        double p[3];
        const uint32_t v0 = std::min(a, b), v1 = std::max(a, b);
        heap.push({placement(v0, v1, p), v0, v1, stamp[v0], stamp[v1]});
    }

    bool collapse(uint32_t v0, uint32_t v1) {
        // link condition: exactly the two apex vertices of the edge's two faces are common neighbours
        std::vector<uint32_t> n0, n1, common;
        neighbours(v0, n0);
        neighbours(v1, n1);
        std::set_intersection(n0.begin(), n0.end(), n1.begin(), n1.end(), std::back_inserter(common));
        uint32_t shared_faces = 0;
        for (uint32_t f : vfaces[v1])
            if (face_alive[f] && (tri[f * 3] == (int32_t)v0 || tri[f * 3 + 1] == (int32_t)v0 || tri[f * 3 + 2] == (int32_t)v0)) shared_faces++;
        if (common.size() != shared_faces || shared_faces == 0 || shared_faces > 2) return false;
        if (n_faces - shared_faces < 4) return false;      // never below a tetrahedron
        double target[3];
        placement(v0, v1, target);
        // normal-flip rejection over the faces that survive
        for (int side = 0; side < 2; side++) {
            const uint32_t v = side ? v1 : v0, other = side ? v0 : v1;
            for (uint32_t f : vfaces[v]) {
                if (!face_alive[f]) continue;
                if (tri[f * 3] == (int32_t)other || tri[f * 3 + 1] == (int32_t)other || tri[f * 3 + 2] == (int32_t)other) continue;
                double before[3], after[3];
                face_normal(f, nullptr, -1, before);
                face_normal(f, target, (int32_t)v, after);
                const double d = before[0] * after[0] + before[1] * after[1] + before[2] * after[2];
                const double la = after[0] * after[0] + after[1] * after[1] + after[2] * after[2];
                if (d <= 0 || la == 0) return false;
            }
        }
        // commit: v1 -> v0
        for (uint32_t f : vfaces[v1]) {
            if (!face_alive[f]) continue;
            bool has_v0 = false;
            for (int k = 0; k < 3; k++) has_v0 |= tri[f * 3 + k] == (int32_t)v0;
            if (has_v0) { face_alive[f] = 0; n_faces--; continue; }
            for (int k = 0; k < 3; k++)
                if (tri[f * 3 + k] == (int32_t)v1) tri[f * 3 + k] = (int32_t)v0;
            vfaces[v0].push_back(f);
        }
        vfaces[v1].clear();
        vfaces[v0].erase(std::remove_if(vfaces[v0].begin(), vfaces[v0].end(), [&](uint32_t f) { return !face_alive[f]; }), vfaces[v0].end());
        vert_alive[v1] = 0;
        pos[(size_t)v0 * 3] = target[0]; pos[(size_t)v0 * 3 + 1] = target[1]; pos[(size_t)v0 * 3 + 2] = target[2];
        Q[v0].add(Q[v1]);
        stamp[v0]++; stamp[v1]++;
        neighbours(v0, n0);
        for (uint32_t w : n0) push(v0, w);
        return true;
    }
};

}  // namespace

extern "C" int mve_mesh_simplify(const float* verts, uint32_t V, const int32_t* faces, uint32_t F, uint32_t target_faces, float* out_verts,
                                 int32_t* out_faces, uint32_t* out_counts) {
    if (!verts || !faces || !out_verts || !out_faces || !out_counts) return -1;
    for (size_t i = 0; i < (size_t)F * 3; i++)
        if (faces[i] < 0 || (uint32_t)faces[i] >= V) return -1;
    Simplifier s;
    s.pos.assign(verts, verts + (size_t)V * 3);
    s.tri.assign(faces, faces + (size_t)F * 3);
    s.face_alive.assign(F, 1);
    s.vert_alive.assign(V, 1);
    s.stamp.assign(V, 0);
    s.Q.assign(V, Quadric());
    s.vfaces.assign(V, {});
    s.n_faces = F;
    std::vector<std::pair<uint32_t, uint32_t>> edges;
    edges.reserve((size_t)F * 3);
    for (uint32_t f = 0; f < F; f++) {
        const int32_t* t = &s.tri[f * 3];
        if (t[0] == t[1] || t[1] == t[2] || t[0] == t[2]) { s.face_alive[f] = 0; s.n_faces--; continue; }
        double n[3];
        s.face_normal(f, nullptr, -1, n);
        const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (len > 0) {
            const double a = n[0] / len, b = n[1] / len, c = n[2] / len;
            const double* p = &s.pos[(size_t)t[0] * 3];
            const double d = -(a * p[0] + b * p[1] + c * p[2]);
            for (int k = 0; k < 3; k++) s.Q[t[k]].add_plane(a, b, c, d, len * 0.5);      // area-weighted
        }
        for (int k = 0; k < 3; k++) {
            s.vfaces[t[k]].push_back(f);
            const uint32_t a = (uint32_t)t[k], b = (uint32_t)t[(k + 1) % 3];
            edges.emplace_back(std::min(a, b), std::max(a, b));
        }
    }
    std::sort(edges.begin(), edges.end());
    edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    for (auto& e : edges) s.push(e.first, e.second);
    while (s.n_faces > target_faces && !s.heap.empty()) {
        const Candidate c = s.heap.top();
        s.heap.pop();
        if (!s.vert_alive[c.v0] || !s.vert_alive[c.v1] || s.stamp[c.v0] != c.s0 || s.stamp[c.v1] != c.s1) continue;
        s.collapse(c.v0, c.v1);                            // a rejected edge is simply dropped until one of its ends changes
    }
    std::vector<int32_t> remap(V, -1);
    uint32_t nv = 0, nf = 0;
    for (uint32_t f = 0; f < F; f++) {
        if (!s.face_alive[f]) continue;
        for (int k = 0; k < 3; k++) {
            const int32_t v = s.tri[f * 3 + k];
            if (remap[v] < 0) {
                remap[v] = (int32_t)nv;
                for (int c = 0; c < 3; c++) out_verts[(size_t)nv * 3 + c] = (float)s.pos[(size_t)v * 3 + c];
                nv++;
            }
            out_faces[(size_t)nf * 3 + k] = remap[v];
        }
        nf++;
    }
    out_counts[0] = nv;
    out_counts[1] = nf;
    return 0;
}
