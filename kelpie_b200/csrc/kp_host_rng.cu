// Host-side replay of the reference's random draws (no device work in this file).
//
// The reference's Kelpie optimizers consume the host generators once per epoch and call
// (pairwise_ranking_optimizer.py:165-195: np.random.shuffle, torch.randint(N+1), torch.randint(2);
// multiclass_nll_optimizer.py:148: torch.randperm).  Both generators are the 32-bit Mersenne Twister
// (torch's CPU default generator; numpy's legacy global RandomState), and every one of those draws is a
// fixed function of consecutive output words:
//   torch.randint(high < 2^32)   word % high                     (serial, one word per element)
//   torch.randperm(n)            Fisher-Yates from the front, z = word % (n - i), n - 1 words
//   np.random.shuffle(1-d)       Fisher-Yates from the back, j = masked rejection sample in [0, i]
// so the 2 * epochs + epochs Python calls of one mimic post-training collapse into one native call each that
// walks the generator state handed in by the host mirror (kelpie_b200/plans.py) and leaves it advanced
// exactly as the reference's calls would.  The mirror checks these against torch / numpy themselves once
// per process and keeps the per-call path where they do not reproduce (another generator layout).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/kelpie_b200.h"

namespace {

constexpr int MT_N = 624, MT_M = 397;

struct Twister {
  uint32_t* key;
  int pos;  // index of the next word; MT_N = regenerate first

  void regenerate() {
    auto mix = [](uint32_t u, uint32_t v) {
      uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
      return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
    };
    int i = 0;
    for (; i < MT_N - MT_M; ++i) key[i] = key[i + MT_M] ^ mix(key[i], key[i + 1]);
    for (; i < MT_N - 1; ++i) key[i] = key[i + MT_M - MT_N] ^ mix(key[i], key[i + 1]);
    key[MT_N - 1] = key[MT_M - 1] ^ mix(key[MT_N - 1], key[0]);
    pos = 0;
  }

  uint32_t next() {
    if (pos >= MT_N) regenerate();
    uint32_t y = key[pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }

  void skip(int64_t count) {
    while (count > 0) {
      if (pos >= MT_N) regenerate();
      int64_t step = count < MT_N - pos ? count : MT_N - pos;
      pos += (int)step;
      count -= step;
    }
  }
};

bool bad_state(const uint32_t* key, const int32_t* pos) { return !key || !pos || *pos < 0 || *pos > MT_N; }

}  // namespace

extern "C" {

int kp_mt19937_words(uint32_t* key, int32_t* pos, int64_t count, uint32_t* out) {
  if (bad_state(key, pos) || count < 0) return KP_EINVAL;
  Twister t{key, *pos};
  if (out) {
    for (int64_t i = 0; i < count; ++i) out[i] = t.next();
  } else {
    t.skip(count);
  }
  *pos = t.pos;
  return KP_OK;
}

int kp_replay_transe_corruptions(uint32_t* key, int32_t* pos, int32_t epochs, int64_t drawn, int64_t used, uint32_t high,
                                 int32_t* neg_code) {
  if (bad_state(key, pos) || epochs < 0 || drawn < 0 || used < 0 || used > drawn || high == 0 || high > 0x80000000u ||
      (!neg_code && (int64_t)epochs * used > 0))
    return KP_EINVAL;
  Twister t{key, *pos};
  for (int32_t e = 0; e < epochs; ++e) {
    int32_t* row = neg_code + (int64_t)e * used;
    for (int64_t i = 0; i < used; ++i) row[i] = (int32_t)(t.next() % high);  // torch.randint(high, (drawn,))[:used]
    t.skip(drawn - used);
    for (int64_t i = 0; i < used; ++i) row[i] = (int32_t)((uint32_t)row[i] | ((t.next() & 1u) << 31));  // randint(2, ...)
    t.skip(drawn - used);
  }
  *pos = t.pos;
  return KP_OK;
}

int kp_replay_numpy_shuffles(uint32_t* key, int32_t* pos, int32_t epochs, int32_t n, int32_t* perm) {
  if (bad_state(key, pos) || epochs < 0 || n < 0 || (!perm && (int64_t)epochs * n > 0)) return KP_EINVAL;
  Twister t{key, *pos};
  for (int32_t e = 0; e < epochs; ++e) {
    int32_t* row = perm + (int64_t)e * n;
    if (e == 0) {
      for (int32_t i = 0; i < n; ++i) row[i] = i;
    } else {
      for (int32_t i = 0; i < n; ++i) row[i] = row[i - n];  // the reference keeps shuffling the same array
    }
    for (int32_t i = n - 1; i >= 1; --i) {
      uint32_t mask = (uint32_t)i;
      mask |= mask >> 1, mask |= mask >> 2, mask |= mask >> 4, mask |= mask >> 8, mask |= mask >> 16;
      uint32_t j;
      while ((j = t.next() & mask) > (uint32_t)i) {
      }
      int32_t tmp = row[i];
      row[i] = row[j];
      row[j] = tmp;
    }
  }
  *pos = t.pos;
  return KP_OK;
}

// One TransE post-training job in one call: the shuffles (numpy's generator) and the corruptions (torch's generator, handed in
// as the 5056-byte buffer of torch.get_rng_state(): seed u64 | left i32 | seeded i32 | next u64 | 624 key words as u64 | normal
// cache) of kp_replay_numpy_shuffles + kp_replay_transe_corruptions, with the positive of training row i of an epoch =
// shuffled row i / ratio (the first n rows of np.repeat(rows, ratio)).  Both generators end as after the separate calls.
int kp_replay_transe_job(uint32_t* np_key, int32_t* np_pos, uint8_t* torch_state, int64_t torch_state_bytes, int32_t epochs, int32_t n,
                         int32_t ratio, uint32_t high, int32_t* pos_idx, int32_t* neg_code) {
  constexpr int64_t T_BYTES = 5056;
  if (bad_state(np_key, np_pos) || !torch_state || torch_state_bytes != T_BYTES || epochs < 0 || n < 0 || ratio < 1 || high == 0 ||
      high > 0x80000000u || ((!pos_idx || !neg_code) && (int64_t)epochs * n > 0))
    return KP_EINVAL;
  int32_t left;
  uint64_t next;
  memcpy(&left, torch_state + 8, 4);
  memcpy(&next, torch_state + 16, 8);
  uint32_t tkey[MT_N];
  for (int i = 0; i < MT_N; ++i) {
    uint64_t w;
    memcpy(&w, torch_state + 24 + 8 * i, 8);
    tkey[i] = (uint32_t)w;
  }
  int32_t tpos = (left == 1 && next == 0) ? MT_N : (int32_t)next;  // freshly seeded: regenerate first
  if (tpos < 0 || tpos > MT_N) return KP_EINVAL;
  Twister nt{np_key, *np_pos}, tt{tkey, tpos};
  std::vector<int32_t> perm((size_t)n);
  for (int32_t i = 0; i < n; ++i) perm[i] = i;
  const int64_t drawn = (int64_t)ratio * n;
  for (int32_t e = 0; e < epochs; ++e) {
    for (int32_t i = n - 1; i >= 1; --i) {  // np.random.shuffle: the reference keeps shuffling the same array
      uint32_t mask = (uint32_t)i;
      mask |= mask >> 1, mask |= mask >> 2, mask |= mask >> 4, mask |= mask >> 8, mask |= mask >> 16;
      uint32_t j;
      while ((j = nt.next() & mask) > (uint32_t)i) {
      }
      const int32_t tmp = perm[i];
      perm[i] = perm[j];
      perm[j] = tmp;
    }
    int32_t* prow = pos_idx + (int64_t)e * n;
    int32_t* crow = neg_code + (int64_t)e * n;
    for (int32_t i = 0; i < n; ++i) prow[i] = perm[i / ratio];
    for (int32_t i = 0; i < n; ++i) crow[i] = (int32_t)(tt.next() % high);  // torch.randint(high, (ratio * n,))[:n]
    tt.skip(drawn - n);
    for (int32_t i = 0; i < n; ++i) crow[i] = (int32_t)((uint32_t)crow[i] | ((tt.next() & 1u) << 31));  // randint(2, ...)
    tt.skip(drawn - n);
  }
  *np_pos = nt.pos;
  for (int i = 0; i < MT_N; ++i) {
    const uint64_t w = tkey[i];
    memcpy(torch_state + 24 + 8 * i, &w, 8);
  }
  next = (uint64_t)tt.pos;
  left = MT_N + 1 - tt.pos;  // torch keeps left + next == 625
  const int32_t seeded = 1;
  memcpy(torch_state + 16, &next, 8);
  memcpy(torch_state + 8, &left, 4);
  memcpy(torch_state + 12, &seeded, 4);
  return KP_OK;
}

}  // extern "C"
