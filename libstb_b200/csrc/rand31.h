/*
 * rand31.h -- glibc's rand() / srand() as a generator with its state in a struct, so that every
 * chain of a batched ARS sampler owns the stream the reference's global generator would give it
 * (the reference's ARMS draws its uniforms as ((double)rand() + 0.5) / 2^31, lib/arms.c:913-918).
 * glibc's default rand() is random_r TYPE_3: the additive feedback generator x^31 + x^3 + 1 over
 * 32-bit words, seeded by the minimal-standard LCG (16807, modulus 2^31 - 1) and warmed up by
 * 310 discarded outputs; an output is the new word shifted right by one bit.
 * tests/test_ars_cpu.py checks bit-equality with libc for several seeds.
 */
#ifndef STB_RAND31_H
#define STB_RAND31_H
#include <stdint.h>

#include "stb_b200.h"

static inline int stb_rand31_step(stb_rand31_t *g) {
  uint32_t v = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
  g->r[g->f] = (int32_t)v;
  if (++g->f >= 31) {
    g->f = 0;
    ++g->b;
  } else if (++g->b >= 31)
    g->b = 0;
  return (int)(v >> 1);
}

static inline void stb_rand31_init(stb_rand31_t *g, unsigned seed) {
  int i;
  int32_t word = seed ? (int32_t)seed : 1;
  g->r[0] = word;
  for (i = 1; i < 31; i++) { /* word = 16807 * word mod (2^31 - 1), Schrage's split */
    const long hi = word / 127773, lo = word % 127773;
    long w = 16807 * lo - 2836 * hi;
    if (w < 0) w += 2147483647;
    word = (int32_t)w;
    g->r[i] = word;
  }
  g->f = 3;
  g->b = 0;
  for (i = 0; i < 310; i++) (void)stb_rand31_step(g);
}

static inline double stb_rand31_unit(void *g) { /* the reference's u_random() on this stream */
  return ((double)stb_rand31_step((stb_rand31_t *)g) + 0.5) / 2147483648.0;
}

#endif
