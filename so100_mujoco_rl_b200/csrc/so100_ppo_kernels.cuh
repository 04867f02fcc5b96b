// so100_ppo_kernels.cuh — fused PPO learner kernels (C ABI: include/so100_ppo.h); included by so100_b200.cu.
//
// What SB3 runs as ~60 small PyTorch kernels per minibatch (policy.evaluate_actions, the losses, loss.backward(),
// clip_grad_norm_, Adam.step; reference call site: stable_baselines3.PPO(...).learn, src/so100_mujoco_rl/main.py:56-64,
// 234-238) is ONE gradient kernel plus two tiny ones here.  The policy is SB3's default MlpPolicy: two separate
// 2x64 tanh towers (pi: od -> 64 -> 64 -> 6, vf: od -> 64 -> 64 -> 1) and a state-independent log_std.
//
// Gradient kernel: persistent CTAs (one per SM, 256 threads), each looping over tiles of 64 samples.  Both towers'
// weights stay in shared memory for the whole kernel (77 KB); a tile's activations live in shared memory in two
// layouts, [feature][sample] for the forward / back-propagation products and [sample][feature] for the weight-gradient
// products, so that every product is a 64x64 register-tiled GEMM (4x4 outputs per thread) fed by 16-byte shared loads
// of which one operand is a warp broadcast.  Weight gradients accumulate in REGISTERS across all tiles of the CTA
// (each thread owns a fixed set of parameters) and leave the SM once, as per-CTA partials; a second kernel adds the
// partials in a fixed order, so the result is deterministic.  fp32 CUDA cores throughout: K = 64 products with exact
// fp32 parity against torch's autograd are the contract here (a tcgen05 TF32 variant is the next step).
#pragma once

namespace ppo {

constexpr int HID = SO100_PPO_HIDDEN, ACT = SO100_PPO_ACT, TB = SO100_PPO_TILE, NT = 256, K1 = 16;
constexpr int LDT = 68;  // row stride of the [sample][feature] copies (16-byte aligned, de-phases the banks)
constexpr int LDX = 20;  // row stride of the [sample][input feature] copy
constexpr float LOG_SQRT_2PI = 0.91893853320467274178f;

struct Layout {  // offsets into the flat parameter vector
  int od, W1[2], b1[2], W2[2], b2[2], W3[2], b3[2], log_std, total;
};
__host__ __device__ inline Layout make_layout(int od) {
  Layout L;
  L.od = od;
  int o = 0;
  for (int t = 0; t < 2; t++) {
    const int nout = t == 0 ? ACT : 1;
    L.W1[t] = o; o += HID * od;
    L.b1[t] = o; o += HID;
    L.W2[t] = o; o += HID * HID;
    L.b2[t] = o; o += HID;
    L.W3[t] = o; o += nout * HID;
    L.b3[t] = o; o += nout;
  }
  L.log_std = o; o += ACT;
  L.total = o;
  return L;
}

// shared-memory image of one tower's weights
struct TowerS {
  float* W1t;  // [K1][HID]   W1t[k][m] = W1[m][k], rows k >= od are zero
  float* W2t;  // [HID][HID]  W2t[k][m] = W2[m][k]
  float* W2n;  // [HID][HID]  W2 as stored, [out][in]
  float* W3;   // [nout][HID]
  float *b1, *b2, *b3;
};
constexpr int tower_floats(bool with_w2n) { return K1 * HID + HID * HID + (with_w2n ? HID * HID : 0) + 8 * HID + HID + HID + 8; }

__device__ inline float* carve_tower(float* p, TowerS& T, bool with_w2n) {
  T.W1t = p; p += K1 * HID;
  T.W2t = p; p += HID * HID;
  T.W2n = with_w2n ? p : nullptr; p += with_w2n ? HID * HID : 0;
  T.W3 = p; p += 8 * HID;
  T.b1 = p; p += HID;
  T.b2 = p; p += HID;
  T.b3 = p; p += 8;
  return p;
}
__device__ inline void load_tower(const Layout& L, const float* P, int t, TowerS& T, bool with_w2n) {
  const int od = L.od, nout = t == 0 ? ACT : 1, tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < K1 * HID; e += nt) { int k = e / HID, m = e % HID; T.W1t[e] = k < od ? P[L.W1[t] + m * od + k] : 0.0f; }
  for (int e = tid; e < HID * HID; e += nt) {
    int k = e / HID, m = e % HID;
    T.W2t[e] = P[L.W2[t] + m * HID + k];
    if (with_w2n) T.W2n[e] = P[L.W2[t] + e];
  }
  for (int e = tid; e < 8 * HID; e += nt) T.W3[e] = e < nout * HID ? P[L.W3[t] + e] : 0.0f;
  for (int e = tid; e < HID; e += nt) { T.b1[e] = P[L.b1[t] + e]; T.b2[e] = P[L.b2[t] + e]; }
  for (int e = tid; e < 8; e += nt) T.b3[e] = e < nout ? P[L.b3[t] + e] : 0.0f;
}

// acc[i][j] += sum_k At[k][r0 + i] * Bm[k][c0 + j]   (both operands K-major in shared memory)
template <int K>
__device__ __forceinline__ void gemm44(const float* __restrict__ At, int lda, const float* __restrict__ Bm, int ldb, int r0, int c0,
                                       float (&acc)[4][4]) {
#pragma unroll 8
  for (int k = 0; k < K; k++) {
    const float4 a = *reinterpret_cast<const float4*>(At + k * lda + r0);
    const float4 b = *reinterpret_cast<const float4*>(Bm + k * ldb + c0);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}
__device__ __forceinline__ void zero44(float (&a)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) a[i][j] = 0.0f;
}
// h = tanh(acc + bias[row]) -> H[row][col] ([feature][sample], ld TB) and, if HT, HT[col][row] ([sample][feature], ld LDT)
__device__ __forceinline__ void store_tanh(const float (&acc)[4][4], const float* bias, int r0, int c0, float* H, float* HT) {
  float h[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) h[i][j] = tanhf(acc[i][j] + bias[r0 + i]);
#pragma unroll
  for (int i = 0; i < 4; i++) *reinterpret_cast<float4*>(H + (r0 + i) * TB + c0) = make_float4(h[i][0], h[i][1], h[i][2], h[i][3]);
  if (HT) {
#pragma unroll
    for (int j = 0; j < 4; j++) *reinterpret_cast<float4*>(HT + (c0 + j) * LDT + r0) = make_float4(h[0][j], h[1][j], h[2][j], h[3][j]);
  }
}
// two hidden layers of one tower on the tile in X ([K1][TB]); leaves H1, H2 (and the transposed copies when given)
__device__ __forceinline__ void tower_forward(const TowerS& T, const float* X, float* H1, float* H1T, float* H2, float* H2T, int r0, int c0) {
  float acc[4][4];
  zero44(acc);
  gemm44<K1>(T.W1t, HID, X, TB, r0, c0, acc);
  store_tanh(acc, T.b1, r0, c0, H1, H1T);
  __syncthreads();
  zero44(acc);
  gemm44<HID>(T.W2t, HID, H1, TB, r0, c0, acc);
  store_tanh(acc, T.b2, r0, c0, H2, H2T);
  __syncthreads();
}
// out[o][s] = b3[o] + sum_k W3[o][k] H2[k][s].  The 64-long sums are split over four thread groups (16 k each, NOUT
// independent chains per thread) and combined through `scratch` ([4][8][TB]); ends with the block synchronised.
template <int NOUT>
__device__ __forceinline__ void head_forward(const TowerS& T, const float* H2, float* out, float* scratch) {
  const int kg = threadIdx.x >> 6, s = threadIdx.x & 63;
  float acc[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; o++) acc[o] = 0.0f;
#pragma unroll
  for (int k = 16 * kg; k < 16 * kg + 16; k++) {
    const float h = H2[k * TB + s];
#pragma unroll
    for (int o = 0; o < NOUT; o++) acc[o] = fmaf(T.W3[o * HID + k], h, acc[o]);
  }
#pragma unroll
  for (int o = 0; o < NOUT; o++) scratch[(kg * 8 + o) * TB + s] = acc[o];
  __syncthreads();
  for (int e = threadIdx.x; e < NOUT * TB; e += NT) {
    const int o = e / TB, c = e % TB;
    out[e] = T.b3[o] + ((scratch[o * TB + c] + scratch[(8 + o) * TB + c]) + (scratch[(16 + o) * TB + c] + scratch[(24 + o) * TB + c]));
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ rollout inference
constexpr int kActSmemFloats = 2 * tower_floats(false) + K1 * TB + 2 * HID * TB + 8 * TB + 32 * TB + 8;

__global__ void __launch_bounds__(NT) act_kernel(Layout L, const float* __restrict__ P, const float* __restrict__ obs, int n, unsigned seed_lo,
                                                 unsigned seed_hi, long long env_offset, unsigned tick, int deterministic, float* act_raw,
                                                 float* act_clip, float* logp, float* value, float* obs_copy) {
  extern __shared__ __align__(16) float sm[];
  TowerS T[2];
  float* p = carve_tower(sm, T[0], false);
  p = carve_tower(p, T[1], false);
  float* X = p; p += K1 * TB;
  float* H1 = p; p += HID * TB;
  float* H2 = p; p += HID * TB;
  float* out = p; p += 8 * TB;
  float* scratch = p; p += 32 * TB;
  float* ls = p;
  const int tid = threadIdx.x, od = L.od;
  load_tower(L, P, 0, T[0], false);
  load_tower(L, P, 1, T[1], false);
  if (tid < ACT) ls[tid] = P[L.log_std + tid];
  const int r0 = 4 * (tid >> 4), c0 = 4 * (tid & 15), ntiles = (n + TB - 1) / TB;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {  // persistent: the weights are staged once per CTA
    const int base = tile * TB;
    __syncthreads();
    for (int e = tid; e < TB * K1; e += NT) {
      const int s = e / K1, f = e % K1, g = base + s;
      float v = 0.0f;
      if (g < n && f < od) {
        v = obs[(size_t)g * od + f];
        if (obs_copy) obs_copy[(size_t)g * od + f] = v;
      }
      X[f * TB + s] = v;
    }
    __syncthreads();
    tower_forward(T[1], X, H1, nullptr, H2, nullptr, r0, c0);
    head_forward<1>(T[1], H2, out + 7 * TB, scratch);  // value in row 7
    tower_forward(T[0], X, H1, nullptr, H2, nullptr, r0, c0);
    head_forward<ACT>(T[0], H2, out, scratch);
    if (tid < TB && base + tid < n) {
      const int g = base + tid;
      float eps[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (!deterministic) {  // 8 standard normals from two Philox blocks (Box-Muller)
#pragma unroll
        for (int b = 0; b < 2; b++) {
          const uint4 r = philox4x32(seed_lo, seed_hi, (unsigned)(env_offset + g), tick, 0x5050u + b, 0u);
          const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const float u1 = ((float)(w[2 * q] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = (float)(w[2 * q + 1] >> 8) * (1.0f / 16777216.0f);
            const float rad = sqrtf(-2.0f * logf(u1));
            float sn, cs;
            sincospif(2.0f * u2, &sn, &cs);
            eps[4 * b + 2 * q] = rad * cs; eps[4 * b + 2 * q + 1] = rad * sn;
          }
        }
      }
      float lp = 0.0f;
#pragma unroll
      for (int k = 0; k < ACT; k++) {
        const float mean = out[k * TB + tid], a = mean + expf(ls[k]) * eps[k];
        lp += -0.5f * eps[k] * eps[k] - ls[k] - LOG_SQRT_2PI;  // (a - mean)^2 / (2 sigma^2) = eps^2 / 2
        if (act_raw) act_raw[(size_t)g * ACT + k] = a;
        if (act_clip) act_clip[(size_t)g * ACT + k] = fminf(fmaxf(a, -1.0f), 1.0f);
      }
      if (logp) logp[g] = lp;
      if (value) value[g] = out[7 * TB + tid];
    }
  }
}

// value tower for single rows, weights from global memory: used where truncations (rare) need V(terminal_obs)
__device__ inline float value_one(const Layout& L, const float* __restrict__ P, const float* __restrict__ x) {
  float h1[HID];
  for (int m = 0; m < HID; m++) {
    float s = P[L.b1[1] + m];
    for (int k = 0; k < L.od; k++) s = fmaf(P[L.W1[1] + m * L.od + k], x[k], s);
    h1[m] = tanhf(s);
  }
  float v = P[L.b3[1]];
  for (int m = 0; m < HID; m++) {
    float s = P[L.b2[1] + m];
    for (int k = 0; k < HID; k++) s = fmaf(P[L.W2[1] + m * HID + k], h1[k], s);
    v = fmaf(P[L.W3[1] + m], tanhf(s), v);
  }
  return v;
}

__global__ void __launch_bounds__(256) post_step_kernel(Layout L, const float* __restrict__ P, int n, const float* __restrict__ reward,
                                                        const uint8_t* __restrict__ term, const uint8_t* __restrict__ trunc,
                                                        const float* __restrict__ tobs, const float* __restrict__ ep_return,
                                                        const int32_t* __restrict__ ep_len, float gamma, float* reward_out, float* done_out,
                                                        double* acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float r = 0.0f, er = 0.0f, el = 0.0f, dn = 0.0f;
  if (i < n) {
    r = reward[i];
    const bool tr = trunc[i] != 0, done = tr || term[i] != 0;
    float ro = r;
    if (tr) ro += gamma * value_one(L, P, tobs + (size_t)i * L.od);  // TimeLimit bootstrap (SB3 on_policy_algorithm.collect_rollouts)
    reward_out[i] = ro;
    done_out[i] = done ? 1.0f : 0.0f;
    if (done) { er = ep_return[i]; el = (float)ep_len[i]; dn = 1.0f; }
  }
  __shared__ float red[4][8];
  float v[4] = {r, er, el, dn};
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += (double)red[threadIdx.x][w];
    if (s != 0.0) atomicAdd(&acc[threadIdx.x], s);
  }
}

__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rew, const float* __restrict__ val, const float* __restrict__ done,
                                                  const float* __restrict__ last_val, int T, int N, float gamma, float lam, float* adv, float* ret) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float last = 0.0f, nxt = last_val[i];
  for (int t = T - 1; t >= 0; t--) {
    const size_t o = (size_t)t * N + i;
    const float nonterm = 1.0f - done[o], v = val[o];
    const float delta = rew[o] + gamma * nxt * nonterm - v;
    last = delta + gamma * lam * nonterm * last;
    adv[o] = last;
    ret[o] = last + v;
    nxt = v;
  }
}

// ------------------------------------------------------------------------------------------------ minibatch gradient
// Per-CTA partial sums (fp64) of adv[idx] and its square over the minibatch; the gradient kernel combines the partials in
// a fixed order (deterministic) into mean and 1 / (unbiased std + 1e-8).
constexpr int kStatCtas = 128;
__global__ void __launch_bounds__(256) adv_stats_kernel(const float* __restrict__ adv, const int64_t* __restrict__ idx, int mb, double* part) {
  __shared__ double red[2][8];
  const int tid = threadIdx.x;
  double s = 0.0, q = 0.0;
  for (int e = blockIdx.x * 256 + tid; e < mb; e += gridDim.x * 256) { const double a = (double)adv[idx[e]]; s += a; q += a * a; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((tid & 31) == 0) { red[0][tid >> 5] = s; red[1][tid >> 5] = q; }
  __syncthreads();
  if (tid < 2) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += red[tid][w];
    part[2 * blockIdx.x + tid] = t;
  }
}

constexpr int kGradSmemFloats = 2 * tower_floats(true)  // weights of both towers
                                + K1 * TB + TB * LDX     // X, XT
                                + 2 * HID * TB           // H1, H2 ([feature][sample]; H2 is overwritten by dZ2)
                                + 4 * TB * LDT           // H1T, H2T, dZ2T, dZ1T ([sample][feature])
                                + 8 * TB + 8 * TB        // head outputs, head gradients
                                + 6 * TB + 4 * TB + 32 * TB + 16;  // actions, (logp_old, adv, ret, valid), head scratch, log_std

__global__ void __launch_bounds__(NT, 1) grad_kernel(Layout L, const float* __restrict__ P, const float* __restrict__ obs, const float* __restrict__ act,
                                                    const float* __restrict__ logp_old, const float* __restrict__ adv, const float* __restrict__ ret,
                                                    const int64_t* __restrict__ idx, int mb, const double* __restrict__ adv_part, float clip,
                                                    float vf_coef, float ent_coef, int normalize, float* __restrict__ gpart) {
  extern __shared__ __align__(16) float sm[];
  TowerS T[2];
  float* p = carve_tower(sm, T[0], true);
  p = carve_tower(p, T[1], true);
  float* X = p; p += K1 * TB;
  float* XT = p; p += TB * LDX;
  float* H1 = p; p += HID * TB;
  float* H2 = p; p += HID * TB;
  float* H1T = p; p += TB * LDT;
  float* H2T = p; p += TB * LDT;
  float* dZ2T = p; p += TB * LDT;
  float* dZ1T = p; p += TB * LDT;
  float* out = p; p += 8 * TB;
  float* dOut = p; p += 8 * TB;
  float* sAct = p; p += 6 * TB;
  float* sOld = p; p += TB;
  float* sAdv = p; p += TB;
  float* sRet = p; p += TB;
  float* sValid = p; p += TB;
  float* scratch = p; p += 32 * TB;
  float* ls = p;

  const int tid = threadIdx.x, od = L.od;
  const int ty = tid >> 4, tx = tid & 15, r0 = 4 * ty, c0 = 4 * tx;
  const int sg = tid >> 6, fk = tid & 63;  // (sample group, feature) mapping of the reductions over a tile's samples
  load_tower(L, P, 0, T[0], true);
  load_tower(L, P, 1, T[1], true);
  if (tid < ACT) ls[tid] = P[L.log_std + tid];
  float a_mean = 0.0f, a_rstd = 1.0f;
  if (normalize) {  // every thread combines the partials in the same order: identical values everywhere
    double s1 = 0.0, s2 = 0.0;
    for (int c = 0; c < kStatCtas; c++) { s1 += adv_part[2 * c]; s2 += adv_part[2 * c + 1]; }
    const double mean = s1 / mb, var = fmax(s2 - s1 * mean, 0.0) / (mb > 1 ? mb - 1 : 1);
    a_mean = (float)mean;
    a_rstd = (float)(1.0 / (sqrt(var) + 1e-8));
  }
  const float inv_mb = 1.0f / (float)mb;

  // gradient accumulators, persistent over the CTA's tiles; each thread owns fixed parameters of both towers
  //   gW2 / gW1: the thread's 4x4 (4x1) block of dW2 (dW1);  gW3p / gb2p / gb1p: partial sums over the thread's 16-sample
  //   group (combined across the 4 groups at the end);  gOut / gLs / loss sums: per sample slot (threads 0..63)
  float gW2[2][4][4], gW1[2][4], gW3p[ACT + 1], gb2p[2] = {0, 0}, gb1p[2] = {0, 0}, gOut[ACT + 1], gLs[ACT];
  float sPg = 0.0f, sV = 0.0f, sKl = 0.0f;
#pragma unroll
  for (int t = 0; t < 2; t++) {
    zero44(gW2[t]);
#pragma unroll
    for (int i = 0; i < 4; i++) gW1[t][i] = 0.0f;
  }
#pragma unroll
  for (int o = 0; o < ACT + 1; o++) { gW3p[o] = 0.0f; gOut[o] = 0.0f; }
#pragma unroll
  for (int o = 0; o < ACT; o++) gLs[o] = 0.0f;

  const int ntiles = (mb + TB - 1) / TB;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();  // the previous tile's buffers are free (also orders the weight loads before first use)
    for (int e = tid; e < TB * K1; e += NT) {
      const int s = e / K1, f = e % K1, g = tile * TB + s;
      float v = 0.0f;
      if (g < mb && f < od) v = obs[(size_t)idx[g] * od + f];
      X[f * TB + s] = v;
      XT[s * LDX + f] = v;
    }
    if (tid < TB) {
      const int g = tile * TB + tid;
      const bool valid = g < mb;
      const size_t src = valid ? (size_t)idx[g] : 0;
#pragma unroll
      for (int k = 0; k < ACT; k++) sAct[k * TB + tid] = valid ? act[src * ACT + k] : 0.0f;
      sOld[tid] = valid ? logp_old[src] : 0.0f;
      sAdv[tid] = valid ? adv[src] : 0.0f;
      sRet[tid] = valid ? ret[src] : 0.0f;
      sValid[tid] = valid ? 1.0f : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 2; t++) {  // unrolled: the accumulators of both towers stay in registers
      const TowerS& W = T[t];
      const int nout = t == 0 ? ACT : 1;
      tower_forward(W, X, H1, H1T, H2, H2T, r0, c0);
      if (t == 0) head_forward<ACT>(W, H2, out, scratch);
      else head_forward<1>(W, H2, out, scratch);
      // ---- loss and its gradient with respect to the head outputs (one thread per sample)
      if (tid < TB) {
        const float valid = sValid[tid];
        if (t == 0) {
          float lp = 0.0f, d[ACT], isig2[ACT];
#pragma unroll
          for (int k = 0; k < ACT; k++) {
            isig2[k] = expf(-2.0f * ls[k]);
            d[k] = sAct[k * TB + tid] - out[k * TB + tid];
            lp += -0.5f * d[k] * d[k] * isig2[k] - ls[k] - LOG_SQRT_2PI;
          }
          const float lr = lp - sOld[tid], ratio = expf(lr);
          const float A = (sAdv[tid] - a_mean) * a_rstd;
          const float lo = 1.0f - clip, hi = 1.0f + clip;
          const float x1 = A * ratio, x2 = A * fminf(fmaxf(ratio, lo), hi);
          const bool inside = ratio >= lo && ratio <= hi;
          // d min(x1, x2) / d lp with torch's tie rule (half to each operand) and clamp's inclusive pass-through
          float g = 0.0f;
          if (x1 < x2) g = x1;
          else if (x1 > x2) g = inside ? x1 : 0.0f;
          else g = 0.5f * x1 + (inside ? 0.5f * x1 : 0.0f);
          const float dlp = -g * inv_mb * valid;
#pragma unroll
          for (int k = 0; k < ACT; k++) {
            const float dm = dlp * d[k] * isig2[k];
            dOut[k * TB + tid] = dm;
            gOut[k] += dm;  // d b3
            gLs[k] += dlp * (d[k] * d[k] * isig2[k] - 1.0f) - ent_coef * inv_mb * valid;
          }
          sPg += -fminf(x1, x2) * valid;
          sKl += ((ratio - 1.0f) - lr) * valid;
        } else {
          const float e = out[tid] - sRet[tid], dv = 2.0f * e * vf_coef * inv_mb * valid;
          dOut[tid] = dv;
          gOut[ACT] += dv;
          sV += e * e * valid;
        }
      }
      __syncthreads();
      // ---- back through the head and the second tanh: dZ2 = (W3^T dOut) * (1 - H2^2), written over H2 and to dZ2T
      {
        float acc[4][4];
        zero44(acc);
        for (int o = 0; o < nout; o++) {
          const float4 w = *reinterpret_cast<const float4*>(W.W3 + o * HID + r0);
          const float4 g = *reinterpret_cast<const float4*>(dOut + o * TB + c0);
          const float wv[4] = {w.x, w.y, w.z, w.w}, gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = fmaf(wv[i], gv[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float4 h = *reinterpret_cast<const float4*>(H2 + (r0 + i) * TB + c0);
          h.x = acc[i][0] * (1.0f - h.x * h.x); h.y = acc[i][1] * (1.0f - h.y * h.y);
          h.z = acc[i][2] * (1.0f - h.z * h.z); h.w = acc[i][3] * (1.0f - h.w * h.w);
          acc[i][0] = h.x; acc[i][1] = h.y; acc[i][2] = h.z; acc[i][3] = h.w;
          *reinterpret_cast<float4*>(H2 + (r0 + i) * TB + c0) = h;  // each thread rewrites only its own 4x4 block
        }
#pragma unroll
        for (int j = 0; j < 4; j++) *reinterpret_cast<float4*>(dZ2T + (c0 + j) * LDT + r0) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      }
      __syncthreads();
      // ---- dZ1 = (W2^T dZ2) * (1 - H1^2) -> dZ1T only (nothing propagates to the observations)
      {
        float acc[4][4];
        zero44(acc);
        gemm44<HID>(W.W2n, HID, H2, TB, r0, c0, acc);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float4 h = *reinterpret_cast<const float4*>(H1 + (r0 + i) * TB + c0);
          acc[i][0] *= 1.0f - h.x * h.x; acc[i][1] *= 1.0f - h.y * h.y; acc[i][2] *= 1.0f - h.z * h.z; acc[i][3] *= 1.0f - h.w * h.w;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) *reinterpret_cast<float4*>(dZ1T + (c0 + j) * LDT + r0) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      }
      __syncthreads();
      // ---- weight gradients (inner dimension = the tile's samples), accumulated in registers
      gemm44<TB>(dZ2T, LDT, H1T, LDT, r0, c0, gW2[t]);  // dW2[out][in] += dZ2[out][s] H1[in][s]
#pragma unroll 8
      for (int s = 0; s < TB; s++) {                     // dW1[out][in = tx] += dZ1[out][s] X[in][s]
        const float4 a = *reinterpret_cast<const float4*>(dZ1T + s * LDT + r0);
        const float x = XT[s * LDX + tx];
        gW1[t][0] = fmaf(a.x, x, gW1[t][0]); gW1[t][1] = fmaf(a.y, x, gW1[t][1]);
        gW1[t][2] = fmaf(a.z, x, gW1[t][2]); gW1[t][3] = fmaf(a.w, x, gW1[t][3]);
      }
      // dW3[o][k = fk] and the hidden biases: partial sums over this thread's 16 samples, NOUT + 2 independent chains
#pragma unroll
      for (int s = 16 * sg; s < 16 * sg + 16; s++) {
        const float h = H2T[s * LDT + fk];
        if (t == 0) {
#pragma unroll
          for (int o = 0; o < ACT; o++) gW3p[o] = fmaf(dOut[o * TB + s], h, gW3p[o]);
        } else {
          gW3p[ACT] = fmaf(dOut[s], h, gW3p[ACT]);
        }
        gb2p[t] += dZ2T[s * LDT + fk];
        gb1p[t] += dZ1T[s * LDT + fk];
      }
      __syncthreads();  // the next tower (or tile) overwrites the activation buffers
    }
  }

  // ---- per-CTA partials: every parameter index is written by exactly one thread
  float* G = gpart + (size_t)blockIdx.x * (L.total + 4);
#pragma unroll
  for (int t = 0; t < 2; t++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
      for (int j = 0; j < 4; j++) G[L.W2[t] + (r0 + i) * HID + c0 + j] = gW2[t][i][j];
      if (tx < od) G[L.W1[t] + (r0 + i) * od + tx] = gW1[t][i];
    }
  }
  __syncthreads();
  float* scr = H1;  // the activation buffers are free now: combine the four sample groups / the 64 sample slots
#pragma unroll
  for (int o = 0; o < ACT + 1; o++) scr[(sg * 8 + o) * HID + fk] = gW3p[o];
#pragma unroll
  for (int t = 0; t < 2; t++) { scr[2048 + (2 * t) * 256 + sg * HID + fk] = gb2p[t]; scr[2048 + (2 * t + 1) * 256 + sg * HID + fk] = gb1p[t]; }
  __syncthreads();
  for (int e = tid; e < (ACT + 1) * HID; e += NT) {
    const int o = e / HID, k = e % HID;
    const float v = (scr[o * HID + k] + scr[(8 + o) * HID + k]) + (scr[(16 + o) * HID + k] + scr[(24 + o) * HID + k]);
    G[(o < ACT ? L.W3[0] + o * HID : L.W3[1]) + k] = v;
  }
  {
    const int which = tid >> 6;  // 0: pi b2, 1: pi b1, 2: vf b2, 3: vf b1
    const float* q = scr + 2048 + which * 256;
    const float v = (q[fk] + q[HID + fk]) + (q[2 * HID + fk] + q[3 * HID + fk]);
    G[(which == 0 ? L.b2[0] : which == 1 ? L.b1[0] : which == 2 ? L.b2[1] : L.b1[1]) + fk] = v;
  }
  __syncthreads();
  if (tid < TB) {
#pragma unroll
    for (int o = 0; o < ACT; o++) { scr[o * TB + tid] = gOut[o]; scr[(ACT + o) * TB + tid] = gLs[o]; }
    scr[12 * TB + tid] = gOut[ACT]; scr[13 * TB + tid] = sPg; scr[14 * TB + tid] = sV; scr[15 * TB + tid] = sKl;
  }
  __syncthreads();
  if (tid < 16) {
    float v = 0.0f;
    for (int k = 0; k < TB; k++) v += scr[tid * TB + k];
    const int dst = tid < ACT ? L.b3[0] + tid : tid < 2 * ACT ? L.log_std + tid - ACT : tid == 12 ? L.b3[1] : L.total + tid - 13;
    G[dst] = v;
  }
}

// grad[i] = sum over CTAs of the partials (fixed order); loss_out = the three loss sums / mb
__global__ void __launch_bounds__(256) reduce_kernel(const float* __restrict__ gpart, int nparts, int total, int mb, float* grad, float* loss_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total + 3) return;
  float s = 0.0f;
  for (int c = 0; c < nparts; c++) s += gpart[(size_t)c * (total + 4) + i];
  if (i < total) grad[i] = s;
  else if (loss_out) loss_out[i - total] = s / (float)mb;
}

// clip_grad_norm_ + Adam, single CTA (the parameter vector has ~10^4 entries)
__global__ void __launch_bounds__(1024) adam_kernel(int n, float* P, const float* __restrict__ grad, float* m, float* v, int32_t* step, float gscale,
                                                    float max_norm, float lr, float b1, float b2, float eps) {
  __shared__ float red[32];
  __shared__ float s_coef;
  const int tid = threadIdx.x;
  float q = 0.0f;
  for (int i = tid; i < n; i += 1024) { const float g = grad[i] * gscale; q = fmaf(g, g, q); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((tid & 31) == 0) red[tid >> 5] = q;
  __syncthreads();
  if (tid == 0) {
    float s = 0.0f;
    for (int w = 0; w < 32; w++) s += red[w];
    const float norm = sqrtf(s);
    s_coef = max_norm > 0.0f ? fminf(max_norm / (norm + 1e-6f), 1.0f) : 1.0f;
    *step += 1;
  }
  __syncthreads();
  const float coef = s_coef * gscale;
  const int t = *step;
  const float bc1 = 1.0f - powf(b1, (float)t), bc2 = 1.0f - powf(b2, (float)t);
  const float step_size = lr / bc1, rsq_bc2 = rsqrtf(bc2);
  for (int i = tid; i < n; i += 1024) {
    const float g = grad[i] * coef;
    const float mi = b1 * m[i] + (1.0f - b1) * g, vi = b2 * v[i] + (1.0f - b2) * g * g;
    m[i] = mi; v[i] = vi;
    P[i] -= step_size * mi / (sqrtf(vi) * rsq_bc2 + eps);
  }
}

}  // namespace ppo

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int so100_ppo_param_count(int obs_dim) {
  if (obs_dim < 1 || obs_dim > ppo::K1) return fail(SO100_ERR_ARG, "obs_dim must be 1..16");
  return ppo::make_layout(obs_dim).total;
}
int64_t so100_ppo_workspace_floats(int obs_dim) {
  const int n = so100_ppo_param_count(obs_dim);
  return n < 0 ? n : (int64_t)SO100_PPO_MAX_CTAS * (n + 4) + 4 * ppo::kStatCtas;
}

static int ppo_smem_optin(const void* fn, int floats) {  // per device; the call is cheap, so it is simply repeated
  CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, floats * 4));
  return SO100_OK;
}

int so100_ppo_act(int obs_dim, const float* params, const float* obs, int n, uint64_t seed, int64_t env_offset, uint32_t tick,
                  int deterministic, float* act_raw, float* act_clip, float* logp, float* value, float* obs_copy, void* stream) {
  if (so100_ppo_param_count(obs_dim) < 0) return SO100_ERR_ARG;
  if (!params || !obs || n <= 0) return fail(SO100_ERR_ARG, "bad argument");
  int rc = ppo_smem_optin((const void*)ppo::act_kernel, ppo::kActSmemFloats);
  if (rc) return rc;
  int dev = 0, sms = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ntiles = (n + ppo::TB - 1) / ppo::TB, grid = ntiles < 2 * sms ? ntiles : 2 * sms;  // 85 KB of shared memory: two CTAs per SM
  ppo::act_kernel<<<grid, ppo::NT, ppo::kActSmemFloats * 4, (cudaStream_t)stream>>>(
      ppo::make_layout(obs_dim), params, obs, n, (unsigned)(seed & 0xFFFFFFFFull), (unsigned)(seed >> 32), env_offset, tick, deterministic,
      act_raw, act_clip, logp, value, obs_copy);
  CU(cudaGetLastError());
  return SO100_OK;
}

int so100_ppo_post_step(int obs_dim, const float* params, int n, const float* reward, const uint8_t* terminated, const uint8_t* truncated,
                        const float* terminal_obs, const float* ep_return, const int32_t* ep_len, float gamma, float* reward_out,
                        float* done_out, double* acc, void* stream) {
  if (so100_ppo_param_count(obs_dim) < 0) return SO100_ERR_ARG;
  if (!params || n <= 0 || !reward || !terminated || !truncated || !terminal_obs || !ep_return || !ep_len || !reward_out || !done_out || !acc)
    return fail(SO100_ERR_ARG, "bad argument");
  ppo::post_step_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ppo::make_layout(obs_dim), params, n, reward, terminated, truncated,
                                                                           terminal_obs, ep_return, ep_len, gamma, reward_out, done_out, acc);
  CU(cudaGetLastError());
  return SO100_OK;
}

int so100_ppo_gae(const float* rew, const float* val, const float* done, const float* last_val, int T, int N, float gamma, float lam,
                  float* adv, float* ret, void* stream) {
  if (!rew || !val || !done || !last_val || !adv || !ret || T <= 0 || N <= 0) return fail(SO100_ERR_ARG, "bad argument");
  ppo::gae_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rew, val, done, last_val, T, N, gamma, lam, adv, ret);
  CU(cudaGetLastError());
  return SO100_OK;
}

int so100_ppo_grad(int obs_dim, const float* params, const float* obs, const float* act, const float* logp_old, const float* adv,
                   const float* ret, const int64_t* idx, int mb, float clip_range, float vf_coef, float ent_coef, int normalize,
                   float* workspace, float* grad, float* loss_out, void* stream) {
  if (so100_ppo_param_count(obs_dim) < 0) return SO100_ERR_ARG;
  if (!params || !obs || !act || !logp_old || !adv || !ret || !idx || mb <= 0 || !workspace || !grad) return fail(SO100_ERR_ARG, "bad argument");
  const ppo::Layout L = ppo::make_layout(obs_dim);
  int rc = ppo_smem_optin((const void*)ppo::grad_kernel, ppo::kGradSmemFloats);
  if (rc) return rc;
  int dev = 0, sms = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ntiles = (mb + ppo::TB - 1) / ppo::TB;
  int grid = ntiles < sms ? ntiles : sms;  // persistent: one CTA per SM
  if (grid > SO100_PPO_MAX_CTAS) grid = SO100_PPO_MAX_CTAS;
  cudaStream_t st = (cudaStream_t)stream;
  // fp64 partials behind the gradient partials; (total + 4) * MAX_CTAS floats is a multiple of 8 bytes
  double* stats = reinterpret_cast<double*>(workspace + (size_t)SO100_PPO_MAX_CTAS * (L.total + 4));
  if (normalize) ppo::adv_stats_kernel<<<ppo::kStatCtas, 256, 0, st>>>(adv, idx, mb, stats);
  ppo::grad_kernel<<<grid, ppo::NT, ppo::kGradSmemFloats * 4, st>>>(L, params, obs, act, logp_old, adv, ret, idx, mb, stats, clip_range, vf_coef,
                                                                   ent_coef, normalize, workspace);
  ppo::reduce_kernel<<<(L.total + 3 + 255) / 256, 256, 0, st>>>(workspace, grid, L.total, mb, grad, loss_out);
  CU(cudaGetLastError());
  return SO100_OK;
}

int so100_ppo_adam(int n_params, float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int32_t* step_count, float grad_scale,
                   float max_grad_norm, float lr, float beta1, float beta2, float eps, void* stream) {
  if (n_params <= 0 || !params || !grad || !exp_avg || !exp_avg_sq || !step_count) return fail(SO100_ERR_ARG, "bad argument");
  ppo::adam_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n_params, params, grad, exp_avg, exp_avg_sq, step_count, grad_scale, max_grad_norm, lr,
                                                          beta1, beta2, eps);
  CU(cudaGetLastError());
  return SO100_OK;
}

}  // extern "C"
