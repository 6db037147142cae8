/*
 * TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the third-party gf2x library.
 *
 * The reference calls gf2x_mul() exactly once, in libbch/bch.c:101, to multiply
 * minimal polynomials into the BCH generator polynomial.  gf2x is not vendored
 * in /root/reference (only external/include/gf2x.h:48-50 declares it) and is not
 * installed in this image.  The product of two polynomials over GF(2) is unique,
 * so a plain shift-and-xor product yields bit-identical generator polynomials.
 *
 * c[0 .. an+bn) = a[0 .. an) * b[0 .. bn)  in GF(2)[x], 64 coefficients per word,
 * bit i of word w = coefficient of x^(64 w + i).
 */
#include <string.h>

void gf2x_mul(unsigned long *c, const unsigned long *a, unsigned long an,
              const unsigned long *b, unsigned long bn)
{
  memset(c, 0, (an + bn) * sizeof(unsigned long));
  for (unsigned long wa = 0; wa < an; wa++) {
    unsigned long aw = a[wa];
    while (aw) {
      const int s = __builtin_ctzl(aw);
      aw &= aw - 1;
      for (unsigned long wb = 0; wb < bn; wb++) {
        c[wa + wb] ^= b[wb] << s;
        if (s) c[wa + wb + 1] ^= b[wb] >> (64 - s);
      }
    }
  }
}
