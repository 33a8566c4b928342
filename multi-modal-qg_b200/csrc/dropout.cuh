// Counter-based inter-layer dropout mask shared by every kernel that applies it (pointwise_bf16.cu,
// lstm_step_tc.cu, lstm_persist.cu): element `idx` of logical stream `sid` is kept iff
// u(seed, sid, idx) >= p.  splitmix64 finaliser over (seed, sid, idx); not ATen's Philox stream.
#pragma once
namespace mmqg {
__device__ __forceinline__ float drop_scale(unsigned long long seed, int sid, unsigned long long idx, float p, float inv_keep) {
  unsigned long long z = seed + (unsigned long long)sid * 0x9E3779B97F4A7C15ull + idx * 0xD1342543DE82EF95ull;
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.f;
}
}  // namespace mmqg
