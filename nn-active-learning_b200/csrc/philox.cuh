// Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
// SC'11).  The MC-dropout masks of the query path (tf.nn.dropout after the layers in model.dropout_layers,
// NN.py:167-171, fed through x_feed_dict = {keep_prob: dropout_rate} at PW_NNAL.py:67-87, 232-282) are a pure
// function of
//     (seed, MC pass, dropout site, GLOBAL pool position, unit index)
// so they do not depend on chunking, on the batch size or on how the pool is sharded over GPUs, and the float64
// oracle (oracle/mc_oracle.py) reproduces them bit for bit.  TensorFlow's own unseeded generator cannot be matched.
//   counter = {unit / 4, pool position, pass, site},  key = {seed low, seed high};  word k of the output decides
//   unit 4 * (unit / 4) + k:  kept  <=>  word < floor(keep_prob * 2^32)
#pragma once
#include <cstdint>

struct DropSpec {
  float keep = 1.f;            // keep probability (model.dropout_rate); kept units are divided by it
  uint32_t thresh = 0;         // floor(keep * 2^32); 0 = dropout off
  uint32_t k0 = 0, k1 = 0;     // seed
  uint32_t pass = 0, site = 0; // MC pass counter, layer index of the dropout site
  long long row0 = 0;          // global pool position of row 0
};

#ifdef __CUDACC__
#define NNAL_HD __host__ __device__ __forceinline__
#else
#define NNAL_HD inline
#endif

NNAL_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
