/* Exhaustive check of the 3-operation division by 17 used by the DTW cost producers (csrc/align.cu:div17_exact):
 *     q0 = x * rc;  r = fma(-17, q0, x);  q = fma(r, rc, q0),   rc = RN(1/17)
 * against the IEEE quotient x / 17, for every normal float32 x (2^31 - 2^24 values, ~45 s on one core), or every
 * `stride`-th one:   gcc -O2 -ffp-contract=off -o div17_check tools/div17_check.c -lm && ./div17_check [stride]
 * Exit code 0 = no mismatch.  (tests/test_oracle_align.py runs a strided pass.) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 10) : 1u;
    const float rc = 1.0f / 17.0f;
    uint32_t rcbits;
    memcpy(&rcbits, &rc, 4);
    if (rcbits != 0x3d70f0f1u) { printf("unexpected RN(1/17) = %08x\n", rcbits); return 2; }
    long bad = 0, n = 0;
    for (uint64_t bits = 0x00800000u; bits < 0x7f800000u; bits += stride) {
        const uint32_t b32 = (uint32_t)bits;
        float x;
        memcpy(&x, &b32, 4);
        for (int sign = 0; sign < 2; ++sign) {
            const float xs = sign ? -x : x;
            const float q0 = xs * rc;
            const float r = fmaf(-17.0f, q0, xs);
            const float q = fmaf(r, rc, q0);
            const float ref = xs / 17.0f;
            if (memcmp(&q, &ref, 4) != 0) {
                if (bad < 10) printf("mismatch x=%a q=%a ref=%a\n", xs, q, ref);
                ++bad;
            }
            ++n;
        }
    }
    printf("checked %ld values (stride %u), mismatches %ld\n", n, stride, bad);
    return bad ? 1 : 0;
}
