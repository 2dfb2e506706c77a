// Device code of K4 shared by g2g_dp.cu (stand-alone) and g2g_fused.cu (K3 + K4 in one launch).
#pragma once
#include "g2g_common.cuh"
#include <math.h>

namespace {


struct DpParams {
  const double* trends;   // [G][4][T]
  const double* cost_in;  // optional [G][T][T]
  uint8_t* align_str;     // [G][2T]
  int32_t* align_len;     // [G]
  double* opt_cost;       // [G]
  uint8_t* l_states;      // [G][(T+1)^2]
  double* cost_out;       // optional
  double* mats_out;       // optional [G][5][(T+1)^2]
  int8_t* bp_out;         // optional
  double* l_out;          // optional
  long long n_genes;
  int n_bins;
  double trans[25];       // I[to][from]
  double start[3];
  double half_n;          // N/2
  double inv_4n;          // 1/(4N)
  double two_c;           // 2*C_model
};

constexpr int kWarpsPerBlock = 4;

constexpr int kMaxPrecomputeBins = 45;   // cost matrix staged in shared memory up to this many bins

__host__ __device__ inline size_t dp_smem_per_warp(int T) {
  size_t b = 0;
  b += (size_t)4 * T * sizeof(double);                    // query-side per-bin terms
  if (T <= kMaxPrecomputeBins) b += (size_t)(4 * T + T * T) * sizeof(double);  // ref-side terms + cost matrix
  b += (size_t)T * T * sizeof(uint16_t);                  // packed back-pointers (5 x 3 bits)
  b += (size_t)(T + 1) * (T + 1);                         // L states
  b += (size_t)2 * T;                                     // reversed alignment string
  return (b + 15) & ~(size_t)15;
}

struct Five {
  double v[5];
};

template <int LW>
__device__ __forceinline__ Five shfl_up5(const Five& a) {
  Five r;
#pragma unroll
  for (int k = 0; k < 5; ++k) r.v[k] = __shfl_up_sync(0xffffffffu, a.v[k], 1, LW);
  return r;
}

// first minimum of five candidates (list.index(min(list)), OrgAlign.py:428-437) as a tournament:
// a later candidate wins only when strictly smaller, so ties keep the lowest index
__device__ __forceinline__ void first_min5(double c0, double c1, double c2, double c3, double c4, double& best,
                                           int& arg) {
  const bool p01 = c1 < c0, p23 = c3 < c2;
  const double m01 = p01 ? c1 : c0, m23 = p23 ? c3 : c2;
  const int a01 = p01 ? 1 : 0, a23 = p23 ? 3 : 2;
  const bool pq = m23 < m01;
  const double m = pq ? m23 : m01;
  const int a = pq ? a23 : a01;
  const bool p4 = c4 < m;
  best = p4 ? c4 : m;
  arg = p4 ? 4 : a;
}
// candidates (src[k] + c) + tr[k], k = 0..4, in the reference's evaluation order (OrgAlign.py:352-374)
__device__ __forceinline__ void best5(const Five& src, double c, const double* tr, double& best, int& arg) {
  first_min5((src.v[0] + c) + tr[0], (src.v[1] + c) + tr[1], (src.v[2] + c) + tr[2], (src.v[3] + c) + tr[3],
             (src.v[4] + c) + tr[4], best, arg);
}
__device__ __forceinline__ void best5_gap(const Five& src, const double* tr, double& best, int& arg) {
  // gap cells cost 0.0 (OrgAlign.py:473-474, 499-500); (x + 0.0) + t == x + t bit for bit
  first_min5(src.v[0] + tr[0], src.v[1] + tr[1], src.v[2] + tr[2], src.v[3] + tr[3], src.v[4] + tr[4], best, arg);
}

// Closed form of the MML match cost (OrgAlign.py:471-503 in the four trend vectors).  Explicit
// roundings: every kernel instantiation (and compiler version) evaluates it with the same
// operations, so dumped costs are the costs the DP used.
struct CostConst { double half_n, two_c, inv_4n; };
__device__ __forceinline__ double match_cost(double t_mu, double t_var, double t_ivar, double t_ln, double s_mu,
                                             double s_var, double s_ivar, double s_ln, const CostConst& k) {
  const double d = __dsub_rn(t_mu, s_mu);
  const double d2 = __dmul_rn(d, d);
  const double a = __fma_rn(__dadd_rn(t_var, d2), s_ivar, __fma_rn(__dadd_rn(s_var, d2), t_ivar, -2.0));
  double u = __fma_rn(k.half_n, a, -k.two_c);
  u = __dadd_rn(__dadd_rn(u, s_ln), t_ln);
  return __dmul_rn(u, k.inv_4n);
}

// FULL = false is the production path (cost from trends, compact outputs only); FULL = true also
// honours cost_in and the dense dump pointers (verification, on-demand per-gene matrices).
// LW lanes serve one gene: 32 in general; for short series (T + 1 <= 16 / 8 columns, B = 1) a warp
// carries 2 / 4 genes in its half / quarter segments so fewer lanes idle.
// The warp (threadIdx.x >> 5) of block blockIdx.x aligns genes [(block * WPB + warp) * GPW, + GPW);
// `smem_raw` is the block's dp_smem_per_warp(T) * WPB * GPW bytes of shared memory.
template <int B, bool FULL, int LW, int WPB = kWarpsPerBlock>
__device__ __forceinline__ void dp_genes(const DpParams& p, unsigned char* smem_raw) {
  constexpr int GPW = 32 / LW;  // genes per warp
  const int T = p.n_bins;
  const int nC = T + 1;
  const int lane = (threadIdx.x & 31) % LW;       // lane inside this gene's segment
  const int sub = (threadIdx.x & 31) / LW;
  const int warp = threadIdx.x >> 5;
  const long long g_first = ((long long)blockIdx.x * WPB + warp) * GPW;
  if (g_first >= p.n_genes) return;  // whole warp leaves together; no block-level barriers below
  const bool active = g_first + sub < p.n_genes;
  const long long g = active ? g_first + sub : g_first;   // idle segments shadow the first gene, stores masked

  unsigned char* base = smem_raw + (size_t)(warp * GPW + sub) * dp_smem_per_warp(T);
  double* q_mu = reinterpret_cast<double*>(base);
  double* q_var = q_mu + T;
  double* q_ivar = q_var + T;
  double* q_ln = q_ivar + T;
  const bool precompute = T <= kMaxPrecomputeBins;
  double* r_terms = q_ln + T;                                    // [4][T] reference-side terms
  double* cost_s = r_terms + 4 * T;                              // [T][T] match cost (query bin, ref bin)
  uint16_t* bp = reinterpret_cast<uint16_t*>(precompute ? cost_s + (size_t)T * T : q_ln + T);
  uint8_t* ls = reinterpret_cast<uint8_t*>(bp + (size_t)T * T);
  uint8_t* sbuf = ls + (size_t)nC * nC;

  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const double* tr = p.trends + (size_t)g * 4 * T;
  const bool have_cost = FULL && p.cost_in != nullptr;
  const double* cin = have_cost ? p.cost_in + (size_t)g * T * T : nullptr;
  double* cout = (FULL && p.cost_out) ? p.cost_out + (size_t)g * T * T : nullptr;
  double* mout = (FULL && p.mats_out) ? p.mats_out + (size_t)g * 5 * nC * nC : nullptr;
  int8_t* bout = (FULL && p.bp_out) ? p.bp_out + (size_t)g * 5 * nC * nC : nullptr;
  double* lout = (FULL && p.l_out) ? p.l_out + (size_t)g * nC * nC : nullptr;

  // ---- prologue: per-bin terms of the closed-form cost -----------------------------------
  for (int t = lane; t < T; t += LW) {
    double mu = tr[2 * T + t], sd = tr[3 * T + t];
    q_mu[t] = mu;
    q_var[t] = sd * sd;
    q_ivar[t] = 1.0 / (sd * sd);
    q_ln[t] = log(sd);
  }
  double s_mu[B], s_var[B], s_ivar[B], s_ln[B];
#pragma unroll
  for (int b = 0; b < B; ++b) {
    int j = lane * B + b;
    bool ok = (j >= 1 && j <= T);
    double mu = ok ? tr[j - 1] : 0.0, sd = ok ? tr[T + j - 1] : 1.0;
    s_mu[b] = mu;
    s_var[b] = sd * sd;
    s_ivar[b] = 1.0 / (sd * sd);
    s_ln[b] = log(sd);
  }

  const CostConst kc = {p.half_n, p.two_c, p.inv_4n};
  if (precompute && !have_cost) {
    // all lanes fill the T x T cost matrix up front: it is off the wavefront's dependent chain
    for (int t = lane; t < T; t += LW) {
      const double mu = tr[t], sd = tr[T + t];
      r_terms[t] = mu;
      r_terms[T + t] = sd * sd;
      r_terms[2 * T + t] = 1.0 / (sd * sd);
      r_terms[3 * T + t] = log(sd);
    }
    __syncwarp();
    const unsigned inv_t = ((1u << 20) + T - 1) / T;   // k / T == (k * inv_t) >> 20 for k < T*T <= 2025
#pragma unroll 4
    for (int k = lane; k < T * T; k += LW) {
      const int qi = (int)(((unsigned)k * inv_t) >> 20), rj = k - qi * T;
      cost_s[k] = match_cost(q_mu[qi], q_var[qi], q_ivar[qi], q_ln[qi], r_terms[rj], r_terms[T + rj],
                             r_terms[2 * T + rj], r_terms[3 * T + rj], kc);
    }
  }
  __syncwarp();

  // ---- row 0 / column 0 initial values (OrgAlign.py:272-331) --------------------------------
  // D[0][j] = (D[0][j-1] + 0.0) + (j==1 ? -ln ProbD : I_dd), sequentially from D[0][0] = 0
  Five up[B];
#pragma unroll
  for (int b = 0; b < B; ++b) {
    int j = lane * B + b;
    if (j == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) up[b].v[k] = 0.0;
    } else {
      double d = 0.0;
      for (int jj = 1; jj <= j && jj <= T; ++jj) d = d + (jj == 1 ? p.start[1] : p.trans[3 * 5 + 3]);
      up[b].v[0] = inf; up[b].v[1] = inf; up[b].v[2] = inf; up[b].v[3] = d; up[b].v[4] = inf;
    }
  }
  // value of row 0 at the column left of this lane's first column (the first "diagonal" input)
  Five nbr_prev;
  {
    int j = lane * B - 1;
    if (j <= 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) nbr_prev.v[k] = 0.0;
    } else {
      double d = 0.0;
      for (int jj = 1; jj <= j; ++jj) d = d + (jj == 1 ? p.start[1] : p.trans[3 * 5 + 3]);
      nbr_prev.v[0] = inf; nbr_prev.v[1] = inf; nbr_prev.v[2] = inf; nbr_prev.v[3] = d; nbr_prev.v[4] = inf;
    }
  }
  // row 0 of the landscape / dumps
#pragma unroll
  for (int b = 0; b < B; ++b) {
    int j = lane * B + b;
    if (j <= T && active) {
      ls[j] = (j == 0) ? 0 : 3;
      if (mout) {
#pragma unroll
        for (int k = 0; k < 5; ++k) mout[(size_t)k * nC * nC + j] = up[b].v[k];
      }
      if (bout) {
#pragma unroll
        for (int k = 0; k < 5; ++k) bout[(size_t)k * nC * nC + j] = (k == 3 && j >= 1) ? 3 : -1;
      }
      if (lout) lout[j] = (j == 0) ? 0.0 : up[b].v[3];
    }
  }
  __syncwarp();

  const int n_lanes = (nC + B - 1) / B;  // lanes that own at least one column
  const int n_steps = T + n_lanes - 1;
  double fin_best = 0.0;  // min over the five matrices at (T, T) and its state, set by the owning lane
  int fin_state = 0;
  // unrolled by two for one or two DP columns per lane: the rotation of the five-state values between steps becomes
  // register renaming (B = 2: ~30 fewer moves per step); with more columns per lane it only costs registers
#pragma unroll (B <= 2 ? 2 : 1)
  for (int s = 0; s < n_steps; ++s) {
    Five nbr = shfl_up5<LW>(up[B - 1]);  // left neighbour's last column, row i (it finished it last step)
    const int i = s - lane + 1;
    if (i >= 1 && i <= T && lane < n_lanes && active) {
      const double t_mu = q_mu[i - 1], t_var = q_var[i - 1], t_ivar = q_ivar[i - 1], t_ln = q_ln[i - 1];
      Five left = nbr, diag = nbr_prev;
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const int j = lane * B + b;
        if (j > T) break;
        Five cur;
        int lstate;
        double lbest;
        if (j == 0) {
          // column 0: only I is finite (OrgAlign.py:287-318)
          cur.v[0] = inf; cur.v[1] = inf; cur.v[2] = inf; cur.v[3] = inf;
          cur.v[4] = up[b].v[4] + (i == 1 ? p.start[0] : p.trans[4 * 5 + 4]);
          lstate = 4;
          lbest = cur.v[4];
          if (bout) {
#pragma unroll
            for (int k = 0; k < 5; ++k) bout[(size_t)k * nC * nC + (size_t)i * nC] = (k == 4) ? 4 : -1;
          }
        } else {
          double c;
          if (have_cost) {
            c = cin[(size_t)(i - 1) * T + (j - 1)];
          } else {
            c = precompute ? cost_s[(size_t)(i - 1) * T + (j - 1)]
                           : match_cost(t_mu, t_var, t_ivar, t_ln, s_mu[b], s_var[b], s_ivar[b], s_ln[b], kc);
          }
          if (cout) cout[(size_t)(i - 1) * T + (j - 1)] = c;
          int am, aw, av, ad, ai;
          if (i == 1 && j == 1) {  // OrgAlign.py:344-350
            cur.v[0] = (diag.v[0] + c) + p.start[2];
            am = 0;
          } else {
            best5(diag, c, &p.trans[0], cur.v[0], am);
          }
          best5(left, c, &p.trans[5], cur.v[1], aw);
          best5(up[b], c, &p.trans[10], cur.v[2], av);
          best5_gap(left, &p.trans[15], cur.v[3], ad);
          best5_gap(up[b], &p.trans[20], cur.v[4], ai);
          G2G_DEVICE_ASSERT(i >= 1 && i <= T && j >= 1 && j <= T && (am | aw | av | ad | ai) < 8);
          bp[(size_t)(i - 1) * T + (j - 1)] = (uint16_t)(am | (aw << 3) | (av << 6) | (ad << 9) | (ai << 12));
          lbest = cur.v[0];
          lstate = 0;
#pragma unroll
          for (int k = 1; k < 5; ++k)
            if (cur.v[k] < lbest) { lbest = cur.v[k]; lstate = k; }
          if (bout) {
            size_t o = (size_t)i * nC + j;
            bout[o] = am; bout[(size_t)nC * nC + o] = aw; bout[(size_t)2 * nC * nC + o] = av;
            bout[(size_t)3 * nC * nC + o] = ad; bout[(size_t)4 * nC * nC + o] = ai;
          }
        }
        ls[(size_t)i * nC + j] = (uint8_t)lstate;
        if (i == T && j == T) { fin_best = lbest; fin_state = lstate; }
        if (mout) {
#pragma unroll
          for (int k = 0; k < 5; ++k) mout[(size_t)k * nC * nC + (size_t)i * nC + j] = cur.v[k];
        }
        if (lout) lout[(size_t)i * nC + j] = lbest;
        diag = up[b];   // (i-1, j) is the diagonal input of column j+1
        left = cur;     // (i, j) is its left input
        up[b] = cur;
      }
    }
    // unconditional: before a lane's first row the neighbour's value IS row 0 of the column to the left (what
    // nbr_prev was initialised with), after its last row nobody reads it; lane 0 never uses either
    nbr_prev = nbr;
  }
  __syncwarp();

  // ---- traceback (OrgAlign.py:539-607) by the lane that owns column T ------------------------
  const int owner = T / B;
  int len = 0;
  if (lane == owner && active) {
    const double best = fin_best;
    int state = fin_state;
    p.opt_cost[g] = best;
    int i = T, j = T;
    const int cap = 2 * T;
    while ((i > 0 || j > 0) && len < cap) {
      int prev;
      if (i == 0) prev = 3;        // row 0: D <- D   (backtrackers_D[0][j] = [0, j-1, 3])
      else if (j == 0) prev = 4;   // column 0: I <- I
      else prev = (bp[(size_t)(i - 1) * T + (j - 1)] >> (3 * state)) & 7;
      if (i == 0) state = 3;
      if (j == 0 && i > 0) state = 4;
      G2G_DEVICE_ASSERT(state >= 0 && state < 5 && i >= 0 && j >= 0);
      sbuf[cap - 1 - len] = (uint8_t)((0x494456574dULL >> (8 * state)) & 0xff);  // "MWVDI"[state]
      ++len;
      if (state == 0) { --i; --j; }
      else if (state == 1 || state == 3) { --j; }
      else { --i; }
      state = prev;
    }
    p.align_len[g] = len;
  }
  len = __shfl_sync(0xffffffffu, len, owner, LW);
  __syncwarp();
  uint8_t* so = p.align_str + (size_t)g * 2 * T;
  if (active) {
#pragma unroll 4
    for (int k = lane; k < 2 * T; k += LW) so[k] = (k < len) ? sbuf[2 * T - len + k] : 0;
    uint8_t* lo = p.l_states + (size_t)g * nC * nC;
#pragma unroll 8
    for (int k = lane; k < nC * nC; k += LW) lo[k] = ls[k];
  }
}

}  // namespace
