// Kernel bodies of the hop-rate front end that more than one launch uses: the whole-utterance kernels of
// nws_encoder.cu and the fused front-end launch of the short-buffer path (nws_noise.cu: nws_front_kernel).
#pragma once
#include "nws_internal.cuh"

// ------------------------------------------------------------------------------------------------
// carry[b][t] = sum over hops t' < t of sum_{n in hop t'} double(f0_up[n])   (generators.py:59:
// torch's CPU cumsum accumulates float32 inputs in double and rounds each output to float32; the
// audio kernel adds the in-hop fp64 prefix to this carry and rounds once).
template <int kCarryThreads>
__device__ __forceinline__ void nws_phase_carry_body(int b, const float* __restrict__ f0, double* __restrict__ carry, int T) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* f = f0 + (size_t)b * T;
  const float inv_hop = (float)T / (float)(T * kHop);
  __shared__ double warp_tot[kCarryThreads / 32];
  __shared__ double chunk_tot;
  double base = 0.0;
  for (int t0 = 0; t0 < T; t0 += kCarryThreads) {
    const int t = t0 + tid;
    double s = 0.0;
    if (t < T) {
      const float fm = f[t > 0 ? t - 1 : 0], fc = f[t], fp = f[t + 1 < T ? t + 1 : T - 1];
      // four independent fp64 chains over the hop's 128 samples (exact for audio-range f0)
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      for (int r = 0; r < kHop; r += 4) {
        double q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const NwsLerp c = nws_lerp_coords(t * kHop + r + u, T, inv_hop);
          const float x0 = c.i0 == t ? fc : (c.i0 < t ? fm : fp);
          const float x1 = c.i1 == t ? fc : (c.i1 < t ? fm : fp);
          q[u] = (double)nws_lerp_apply(c, x0, x1);
        }
        s0 += q[0]; s1 += q[1]; s2 += q[2]; s3 += q[3];
      }
      s = (s0 + s1) + (s2 + s3);
    }
    double v = s;  // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    double pre = 0.0;
    for (int w = 0; w < warp; ++w) pre += warp_tot[w];
    if (t < T) carry[(size_t)b * T + t] = base + pre + (v - s);
    if (tid == kCarryThreads - 1) chunk_tot = pre + v;
    __syncthreads();
    base += chunk_tot;
    __syncthreads();
  }
}


// ------------------------------------------------------------------------------------------------
// GRU(2 -> 128), h0 = 0, gate order r,z,n (neural_waveshaping.py:21,25).
// One CTA per utterance, 384 threads = one per gate row; the recurrence is the only sequential dependency at
// hop rate.  A warp owns 32 rows.  The dot products W_hh[row] . h are computed K-SPLIT across the warp: lane l
// holds the four columns k = 4l..4l+3 of all 32 rows of its warp in registers (128 floats, for all T steps),
// reads only its own four h values per step (one conflict-free LDS.128 per warp instead of 32 broadcast
// loads), forms 32 four-term partial sums and the warp combines them with a reduce-scatter of xor-shuffles.
// Register j of lane l accumulates row (j ^ l): with that swizzle every exchange stage is
// `v[j] += shfl_xor(v[j | o], o)` with static register indices and no selects, and lane l ends up with the
// complete dot product of row l.
__device__ __forceinline__ float nws_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void
nws_gru_body(int b, const float* __restrict__ w_hh, const float* __restrict__ w_ih, const float* __restrict__ b_ih,
               const float* __restrict__ b_hh, const float* __restrict__ control, int ctrl_channels,
               float* __restrict__ hbuf, int T, int t_begin, int t_end, float* __restrict__ h_state) {
  const int r = threadIdx.x, lane = r & 31, row0 = r & ~31;
  __shared__ __align__(16) float h_s[2][kEmb];
  __shared__ float pre_rz[2 * kEmb];
  __shared__ float pre_ni[kEmb], pre_nh[kEmb];

  // rows j and j|16 are kept as packed pairs so one fma.rn.f32x2 (FFMA2) advances both dot products: 64 packed
  // FMAs per step instead of 128 scalar ones; each element still goes through the same x, y, z, w fma chain
  float2 wp[16][4];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 lo = *reinterpret_cast<const float4*>(w_hh + (size_t)(row0 + (j ^ lane)) * kEmb + 4 * lane);
    const float4 hi = *reinterpret_cast<const float4*>(w_hh + (size_t)(row0 + ((j | 16) ^ lane)) * kEmb + 4 * lane);
    wp[j][0] = make_float2(lo.x, hi.x); wp[j][1] = make_float2(lo.y, hi.y);
    wp[j][2] = make_float2(lo.z, hi.z); wp[j][3] = make_float2(lo.w, hi.w);
  }
  const float wi0 = w_ih[r * 2], wi1 = w_ih[r * 2 + 1], bi = b_ih[r], bh = b_hh[r];
  const float* c0 = control + (size_t)b * ctrl_channels * T;
  const float* c1 = c0 + T;
  // steps [t_begin, t_end): the recurrence can be run in time blocks (h carried through h_state) so that the
  // rest of the forward can start on the frames that are already encoded
  if (r < kEmb) h_s[t_begin & 1][r] = t_begin > 0 ? h_state[(size_t)b * kEmb + r] : 0.0f;
  float x0 = c0[t_begin], x1 = c1[t_begin];
  __syncthreads();

  for (int t = t_begin; t < t_end; ++t) {
    const float* h = h_s[t & 1];
    const float nx0 = t + 1 < t_end ? c0[t + 1] : 0.0f, nx1 = t + 1 < t_end ? c1[t + 1] : 0.0f;  // prefetch
    const float4 hv = *reinterpret_cast<const float4*>(h + 4 * lane);
    float v[16];
#pragma unroll
    const float2 hx = make_float2(hv.x, hv.x), hy = make_float2(hv.y, hv.y), hz = make_float2(hv.z, hv.z), hw = make_float2(hv.w, hv.w);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float2 ac = __fmul2_rn(wp[j][0], hx);   // (.x: row j ^ lane, .y: row (j | 16) ^ lane)
      ac = __ffma2_rn(wp[j][1], hy, ac);
      ac = __ffma2_rn(wp[j][2], hz, ac);
      ac = __ffma2_rn(wp[j][3], hw, ac);
      v[j] = ac.x + __shfl_xor_sync(0xffffffffu, ac.y, 16);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j | 8], 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j | 4], 4);
#pragma unroll
    for (int j = 0; j < 2; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j | 2], 2);
    const float gh = (v[0] + __shfl_xor_sync(0xffffffffu, v[1], 1)) + bh;   // full dot product of row r
    const float gi = fmaf(wi1, x1, fmaf(wi0, x0, bi));
    if (r < 2 * kEmb) {
      pre_rz[r] = nws_sigmoid(gi + gh);   // r and z gates: activated by the 256 threads that own their rows
    } else {
      pre_ni[r - 2 * kEmb] = gi;
      pre_nh[r - 2 * kEmb] = gh;
    }
    __syncthreads();
    if (r < kEmb) {
      const float rg = pre_rz[r];
      const float zg = pre_rz[kEmb + r];
      const float ng = tanhf(fmaf(rg, pre_nh[r], pre_ni[r]));
      const float hn = fmaf(zg, h[r] - ng, ng);  // (1-z)*n + z*h
      h_s[(t + 1) & 1][r] = hn;
      hbuf[((size_t)b * T + t) * kEmb + r] = hn;
    }
    x0 = nx0; x1 = nx1;
    __syncthreads();
  }
  if (h_state && r < kEmb) h_state[(size_t)b * kEmb + r] = h_s[t_end & 1][r];
}

