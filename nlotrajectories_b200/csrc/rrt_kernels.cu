// Initial guesses on the device: the RRT tree search of core/trajectory_initialization.py:175-216 for P planners at
// once, one WARP per planner running its whole search inside one kernel, against the exact obstacle SDF of the YAML's scene
// (circles, squares, polygons / trapezoids / elliptical half-rings as polygons: core/sdf/casadi.py:27-45, 48-66, 135-148, 218-246;
// union = min, core/sdf/casadi.py:381-383).
//
// Per iteration, per planner (the lock-step tensor program of rrt_device.py, which this kernel reproduces draw for draw):
//   sample     goal with probability goal_sample_rate, else uniform in the bounds; the uniforms are a counter hash of
//              (seed of the start, iteration, draw), so a start depends on its own seed only
//   nearest    lanes scan the tree nodes, warp arg-min (lowest index wins ties, like torch.argmin)
//   steer      step_size from the nearest node towards the sample
//   collision  both ends and the midpoint of the new edge keep `inflation` of clearance: polygon edges are spread over the lanes
//   grow       append the node; within step_size of the goal: append the goal and stop
// fp64 throughout (the trees decide by comparisons; the path post-processing on the host is fp64 too).  The paths (goal -> root, reversed)
// are extracted by the same warp, so only the paths travel back, not the trees.
#include "nlo_common.cuh"
#include <vector>

namespace {

struct RrtScene {                  // passed by value
  const nlo_rrt_obstacle* obs;     // device
  const double* verts;             // device: [n][2]
  int n_obs;
  double sx, sy, gx, gy, lox, loy, hix, hiy, step, inflation, goal_rate;
  int max_iter, M, max_path, post;
};

__device__ __forceinline__ double rrt_uniform(long long seed, long long it, long long k) {
  const long long M31 = (1ll << 31) - 1;
  long long x = (seed * 1103515245ll + (it * 40503ll + k * 9973ll + 12345ll)) & M31;
  const long long mult[3] = {1664525ll, 22695477ll, 1103515245ll}, add[3] = {1013904223ll, 1ll, 12345ll};
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    x = (x ^ (x >> 15)) & M31;
    x = (x * mult[r] + add[r]) & M31;
  }
  x = (x ^ (x >> 13)) & M31;
  return (double)x / 2147483648.0;
}

// exact scene SDF at one point, evaluated by the whole warp (every lane returns the value)
__device__ double rrt_scene_sdf(const RrtScene& S, double px, double py, int lane) {
  double best = HUGE_VAL;
  for (int o = 0; o < S.n_obs; ++o) {
    const nlo_rrt_obstacle ob = S.obs[o];
    double v;
    if (ob.kind == NLO_RRT_CIRCLE) {
      const double dx = px - ob.cx, dy = py - ob.cy;
      v = sqrt(dx * dx + dy * dy) - (ob.size + ob.margin);
    } else if (ob.kind == NLO_RRT_SQUARE) {
      const double half = ob.size * 0.5 + ob.margin;
      const double dx = fabs(px - ob.cx) - half, dy = fabs(py - ob.cy) - half;
      const double ox = dx > 0.0 ? dx : 0.0, oy = dy > 0.0 ? dy : 0.0;
      const double mx = dx > dy ? dx : dy;
      v = sqrt(ox * ox + oy * oy) + (mx < 0.0 ? mx : 0.0);
    } else {                                                     // polygon: edges over the lanes
      double dmin = HUGE_VAL;
      int cross = 0;
      const int nv = (int)ob.n_vertices;
      for (int e = lane; e < nv; e += 32) {
        const double ax = S.verts[2 * (ob.first_vertex + e)], ay = S.verts[2 * (ob.first_vertex + e) + 1];
        const int e1 = e + 1 == nv ? 0 : e + 1;
        const double bx = S.verts[2 * (ob.first_vertex + e1)], by = S.verts[2 * (ob.first_vertex + e1) + 1];
        const double ex = bx - ax, ey = by - ay, wx = px - ax, wy = py - ay;
        double den = ex * ex + ey * ey;
        if (den < 1e-300) den = 1e-300;
        double t = (wx * ex + wy * ey) / den;
        t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
        const double rx = wx - t * ex, ry = wy - t * ey, d2 = rx * rx + ry * ry;
        dmin = d2 < dmin ? d2 : dmin;
        if ((ay > py) != (by > py)) {
          const double xint = ax + (py - ay) * ex / (ey == 0.0 ? 1.0 : ey);
          if (px < xint) ++cross;
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, dmin, off);
        dmin = od < dmin ? od : dmin;
        cross += __shfl_xor_sync(0xffffffffu, cross, off);
      }
      const double d = sqrt(dmin);
      v = ((cross & 1) ? -d : d) - ob.margin;
    }
    best = v < best ? v : best;
  }
  return best;
}

// every ~step along the segment p1 -> p2 the scene keeps `inflation` of clearance (trajectory_initialization.py:_collision_free);
// evaluated by the whole warp, every lane returns the verdict
__device__ bool rrt_segment_free(const RrtScene& S, double x1, double y1, double x2, double y2, int lane) {
  const double dx = x2 - x1, dy = y2 - y1;
  int n = (int)ceil(sqrt(dx * dx + dy * dy) / S.step);
  if (n < 1) n = 1;
  for (int q = 0; q <= n; ++q) {
    const double t = (double)q / (double)n;
    if (!(rrt_scene_sdf(S, x1 + dx * t, y1 + dy * t, lane) >= S.inflation)) return false;
  }
  return true;
}

__global__ void __launch_bounds__(128) rrt_trees_kernel(RrtScene S, const long long* __restrict__ seeds, size_t P, double* __restrict__ pos,
                                                        int* __restrict__ parent, double* __restrict__ path, int* __restrict__ path_len) {
  const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= P) return;
  double* X = pos + wid * (size_t)S.M * 2;
  int* Pa = parent + wid * (size_t)S.M;
  const long long seed = seeds[wid];
  if (lane == 0) { X[0] = S.sx; X[1] = S.sy; Pa[0] = -1; }
  __syncwarp();
  int n_nodes = 1, final_node = -1;
  for (int it = 0; it < S.max_iter && final_node < 0; ++it) {
    const bool to_goal = rrt_uniform(seed, it, 0) < S.goal_rate;
    const double rx = to_goal ? S.gx : S.lox + (S.hix - S.lox) * rrt_uniform(seed, it, 1);
    const double ry = to_goal ? S.gy : S.loy + (S.hiy - S.loy) * rrt_uniform(seed, it, 2);
    // nearest node: arg-min of the squared distance, lowest index on ties
    double bd = HUGE_VAL;
    int bi = 0x7fffffff;
    for (int j = lane; j < n_nodes; j += 32) {
      const double dx = X[2 * j] - rx, dy = X[2 * j + 1] - ry, d2 = dx * dx + dy * dy;
      if (d2 < bd) { bd = d2; bi = j; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    const double nx0 = X[2 * bi], ny0 = X[2 * bi + 1];
    const double dirx = rx - nx0, diry = ry - ny0, norm = sqrt(dirx * dirx + diry * diry);
    const double inv = 1.0 / (norm > 1e-300 ? norm : 1e-300);
    const double newx = nx0 + dirx * inv * S.step, newy = ny0 + diry * inv * S.step;
    bool free_edge = true;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double tq = 0.5 * q;
      free_edge = free_edge && (rrt_scene_sdf(S, nx0 + (newx - nx0) * tq, ny0 + (newy - ny0) * tq, lane) >= S.inflation);
    }
    if (free_edge && norm > 0.0) {
      if (lane == 0) { X[2 * n_nodes] = newx; X[2 * n_nodes + 1] = newy; Pa[n_nodes] = bi; }
      ++n_nodes;
      const double gdx = newx - S.gx, gdy = newy - S.gy;
      if (sqrt(gdx * gdx + gdy * gdy) < S.step) {
        if (lane == 0) { X[2 * n_nodes] = S.gx; X[2 * n_nodes + 1] = S.gy; Pa[n_nodes] = n_nodes - 1; }
        final_node = n_nodes;
        ++n_nodes;
      }
      __syncwarp();
    }
  }
  // path: walk the parents from the goal node to the root, then store it root first
  double* out = path + wid * (size_t)S.max_path * 2;
  int len = -1;
  if (final_node >= 0) {
    len = 0;
    for (int node = final_node; node >= 0; node = Pa[node]) ++len;
    if (len > S.max_path) len = -2;                              // longer than the caller's buffer
    else if (lane == 0) {
      int k = len - 1;
      for (int node = final_node; node >= 0; node = Pa[node], --k) { out[2 * k] = X[2 * node]; out[2 * k + 1] = X[2 * node + 1]; }
    }
  }
  __syncwarp();
  if (S.post && len > 2) {
    // (1) a midpoint before every corner sharper than 60 degrees (insert_intermediate_points), into the tree array as scratch
    double* T = X;
    int m = 0;
    if (2 * len > S.M) len = -2;
    else {
      if (lane == 0) {
        T[0] = out[0]; T[1] = out[1]; m = 1;
        for (int i = 1; i < len - 1; ++i) {
          const double v1x = out[2 * i] - out[2 * i - 2], v1y = out[2 * i + 1] - out[2 * i - 1];
          const double v2x = out[2 * i + 2] - out[2 * i], v2y = out[2 * i + 3] - out[2 * i + 1];
          double c = (v1x * v2x + v1y * v2y) / (sqrt(v1x * v1x + v1y * v1y) * sqrt(v2x * v2x + v2y * v2y));
          c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
          if (acos(c) * 57.29577951308232 > 60.0) { T[2 * m] = (out[2 * i] + out[2 * i - 2]) / 2; T[2 * m + 1] = (out[2 * i + 1] + out[2 * i - 1]) / 2; ++m; }
          T[2 * m] = out[2 * i]; T[2 * m + 1] = out[2 * i + 1]; ++m;
        }
        T[2 * m] = out[2 * len - 2]; T[2 * m + 1] = out[2 * len - 1]; ++m;
      }
      m = __shfl_sync(0xffffffffu, m, 0);
      __syncwarp();
      // (2) greedy shortcut: from node i jump to the farthest node the straight segment to which is collision free (_shortcut)
      int n_out = 1, i = 0;                                     // out[0] is the start already
      while (i < m - 1) {
        int j = m - 1;
        while (j > i + 1 && !rrt_segment_free(S, T[2 * i], T[2 * i + 1], T[2 * j], T[2 * j + 1], lane)) --j;
        if (n_out >= S.max_path) { n_out = -2; break; }          // longer than the caller's buffer
        if (lane == 0) { out[2 * n_out] = T[2 * j]; out[2 * n_out + 1] = T[2 * j + 1]; }
        ++n_out;
        i = j;
      }
      len = n_out;
    }
  }
  if (lane == 0) path_len[wid] = len;
}

}  // namespace

extern "C" int nlo_rrt_paths(const nlo_rrt_obstacle* obs, int n_obs, const double* vertices, int n_vertices, const double* start, const double* goal,
                             const double* lo, const double* hi, const long long* seeds, size_t P, double step_size, int max_iter,
                             double inflation, double goal_sample_rate, int max_path, int postprocess, int device, double* path_host,
                             int* path_len_host) {
  if (!obs || n_obs < 1 || !start || !goal || !lo || !hi || !seeds || !path_host || !path_len_host) return nlo_fail("null argument");
  if (P == 0) return 0;
  if (max_iter < 1 || max_path < 2 || !(step_size > 0.0)) return nlo_fail("rrt: bad parameters");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return nlo_fail("no CUDA device available: libnlo_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return nlo_fail("device %d out of range (have %d)", device, ndev);
  for (int o = 0; o < n_obs; ++o) {
    if (obs[o].kind > NLO_RRT_POLYGON) return nlo_fail("rrt: unknown obstacle kind %u", obs[o].kind);
    if (obs[o].kind == NLO_RRT_POLYGON && (!vertices || obs[o].n_vertices < 3 || (int)(obs[o].first_vertex + obs[o].n_vertices) > n_vertices))
      return nlo_fail("rrt: polygon %d has a bad vertex range", o);
  }
  NLO_CUDA(cudaSetDevice(device));
  RrtScene S;
  S.n_obs = n_obs; S.sx = start[0]; S.sy = start[1]; S.gx = goal[0]; S.gy = goal[1]; S.lox = lo[0]; S.loy = lo[1]; S.hix = hi[0]; S.hiy = hi[1];
  S.step = step_size; S.inflation = inflation; S.goal_rate = goal_sample_rate; S.max_iter = max_iter; S.M = max_iter + 3; S.max_path = max_path; S.post = postprocess;
  nlo_rrt_obstacle* d_obs = nullptr; double* d_verts = nullptr; long long* d_seeds = nullptr; double* d_pos = nullptr; int* d_parent = nullptr;
  double* d_path = nullptr; int* d_len = nullptr;
  auto cleanup = [&]() { for (void* b : {(void*)d_obs, (void*)d_verts, (void*)d_seeds, (void*)d_pos, (void*)d_parent, (void*)d_path, (void*)d_len}) if (b) cudaFree(b); };
  const size_t nv = n_vertices > 0 ? (size_t)n_vertices : 1;
  bool ok = cudaMalloc(&d_obs, n_obs * sizeof(nlo_rrt_obstacle)) == cudaSuccess && cudaMalloc(&d_verts, nv * 2 * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&d_seeds, P * sizeof(long long)) == cudaSuccess && cudaMalloc(&d_pos, P * (size_t)S.M * 2 * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&d_parent, P * (size_t)S.M * sizeof(int)) == cudaSuccess && cudaMalloc(&d_path, P * (size_t)max_path * 2 * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&d_len, P * sizeof(int)) == cudaSuccess;
  if (!ok) { cleanup(); return nlo_fail("rrt: device allocation failed (%zu planners x %d nodes)", P, S.M); }
  ok = cudaMemcpy(d_obs, obs, n_obs * sizeof(nlo_rrt_obstacle), cudaMemcpyHostToDevice) == cudaSuccess &&
       (n_vertices <= 0 || cudaMemcpy(d_verts, vertices, (size_t)n_vertices * 2 * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess) &&
       cudaMemcpy(d_seeds, seeds, P * sizeof(long long), cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) { cleanup(); return nlo_fail("rrt: upload failed"); }
  S.obs = d_obs; S.verts = d_verts;
  rrt_trees_kernel<<<(unsigned)((P * 32 + 127) / 128), 128>>>(S, d_seeds, P, d_pos, d_parent, d_path, d_len);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) { nlo_count_launch(); e = cudaDeviceSynchronize(); }
  if (e == cudaSuccess) e = cudaMemcpy(path_host, d_path, P * (size_t)max_path * 2 * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(path_len_host, d_len, P * sizeof(int), cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return nlo_fail("rrt: %s", cudaGetErrorString(e));
  return 0;
}
