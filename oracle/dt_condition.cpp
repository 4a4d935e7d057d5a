// dt_condition.cpp -- host-side synthetic DEM generator and priority-flood+epsilon.
//
// TEST / BENCH INFRASTRUCTURE ONLY (see dt_oracle.c header).  The reference has no
// DEM conditioning (its fixtures were filled by an external GIS, Example/example.py:33-39);
// SURVEY.md 8(d) freezes the synthetic recipe used by the benchmark and this file is its
// CPU statement.  The product library has device kernels (dtb_synth_dem_f32,
// dtb_fill_depressions_f32) that must reproduce these outputs bit-for-bit; the tests
// compare them.
//
// Recipe "dtb-synth-v1" (all arithmetic in IEEE f32, one rounding per operation, no FMA):
//   z(r,c) = z0 + sr*r + sc*c + sum_k amp[k]*vnoise(r,c; L_k = 2048>>k, k=0..9)
//            - depth * ridge(r,c)^2,  ridge = max(0, 1 - 8*|vnoise(r,c; L=1024, oct=31)|)
// vnoise = bilinear value noise with smoothstep weights on an integer-hashed lattice.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <queue>
#include <vector>

namespace {

inline uint32_t lattice_hash(uint32_t ix, uint32_t iy, uint32_t oct, uint32_t seed)
{
    uint32_t h = seed ^ (ix * 0x9E3779B1u) ^ (iy * 0x85EBCA77u) ^ (oct * 0xC2B2AE3Du);
    h ^= h >> 16; h *= 0x7FEB352Du;
    h ^= h >> 15; h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}

inline float lattice_val(uint32_t ix, uint32_t iy, uint32_t oct, uint32_t seed)
{
    return (float)(lattice_hash(ix, iy, oct, seed) >> 8) * (1.0f / 8388608.0f) - 1.0f;
}

inline float smooth(float t)
{
    const float t2 = t * t;
    const float b = 3.0f - 2.0f * t;
    return t2 * b;
}

inline float vnoise(int64_t r, int64_t c, int L, uint32_t oct, uint32_t seed)
{
    const uint32_t ix = (uint32_t)(c / L), iy = (uint32_t)(r / L);
    const float fx = (float)(c % L) / (float)L, fy = (float)(r % L) / (float)L;
    const float sx = smooth(fx), sy = smooth(fy);
    const float v00 = lattice_val(ix, iy, oct, seed), v10 = lattice_val(ix + 1, iy, oct, seed);
    const float v01 = lattice_val(ix, iy + 1, oct, seed), v11 = lattice_val(ix + 1, iy + 1, oct, seed);
    const float d0 = v10 - v00, d1 = v11 - v01;
    const float a = v00 + sx * d0;
    const float b = v01 + sx * d1;
    const float d = b - a;
    return a + sy * d;
}

inline float succ(float x)
{
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if (x > 0.0f) u += 1;
    else if (x < 0.0f) u -= 1;
    else u = 1u; // +-0 -> smallest positive
    std::memcpy(&x, &u, 4);
    return x;
}

} // namespace

extern "C" {

// amp[10]: octave amplitudes; params: z0, sr, sc, depth.  Rows [row0, row0+rows) of the
// conceptual global raster are written to out[rows*cols].
void orc_synth_dem_f32(int64_t rows, int64_t cols, int64_t row0, uint32_t seed, const float *amp,
                       float z0, float sr, float sc, float depth, float *out)
{
    #pragma omp parallel for schedule(static)
    for (int64_t rl = 0; rl < rows; ++rl) {
        const int64_t r = row0 + rl;
        for (int64_t c = 0; c < cols; ++c) {
            float z = z0 + sr * (float)r;
            z = z + sc * (float)c;
            for (int k = 0; k < 10; ++k) {
                const float v = vnoise(r, c, 2048 >> k, (uint32_t)k, seed);
                z = z + amp[k] * v;
            }
            const float n2 = vnoise(r, c, 1024, 31u, seed);
            float ridge = 1.0f - 8.0f * std::fabs(n2);
            if (ridge < 0.0f) ridge = 0.0f;
            z = z - depth * (ridge * ridge);
            out[rl * cols + c] = z;
        }
    }
}

// Priority-flood + epsilon (Barnes, Lehman & Mulla 2014, Alg. 3 semantics): every cell
// that is not a seed ends up strictly above at least one neighbour.  Seeds = valid cells
// on the raster edge or adjacent to a nodata cell (== -100 or NaN).  In place.
// The result is the unique fixed point of W(c) = max(z(c), min_n succ(W(n))).
void orc_priority_flood_eps_f32(float *dem, int64_t rows, int64_t cols)
{
    const int64_t n = rows * cols;
    std::vector<uint8_t> closed((size_t)n, 0);
    typedef std::pair<float, int64_t> item;
    std::priority_queue<item, std::vector<item>, std::greater<item>> pq;
    auto is_nd = [&](int64_t p) { return dem[p] == -100.0f || dem[p] != dem[p]; };
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < cols; ++c) {
            const int64_t p = r * cols + c;
            if (is_nd(p)) { closed[p] = 1; continue; }
            bool seed = (r == 0 || c == 0 || r == rows - 1 || c == cols - 1);
            for (int dr = -1; dr <= 1 && !seed; ++dr)
                for (int dc = -1; dc <= 1 && !seed; ++dc) {
                    const int64_t rr = r + dr, cc = c + dc;
                    if (rr < 0 || rr >= rows || cc < 0 || cc >= cols) continue;
                    if (is_nd(rr * cols + cc)) seed = true;
                }
            if (seed) { closed[p] = 1; pq.push(item(dem[p], p)); }
        }
    while (!pq.empty()) {
        const item it = pq.top();
        pq.pop();
        const int64_t p = it.second, r = p / cols, c = p % cols;
        const float up = succ(it.first);
        for (int dr = -1; dr <= 1; ++dr)
            for (int dc = -1; dc <= 1; ++dc) {
                const int64_t rr = r + dr, cc = c + dc;
                if (rr < 0 || rr >= rows || cc < 0 || cc >= cols) continue;
                const int64_t q = rr * cols + cc;
                if (closed[q]) continue;
                closed[q] = 1;
                if (dem[q] < up) dem[q] = up;
                pq.push(item(dem[q], q));
            }
    }
}

} // extern "C"
