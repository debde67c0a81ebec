/* Exhaustive check of the identity behind div3_exact() in csrc/vsc_kernels.cuh: for every float x in [0, 2295 + a few
 * ulps] (2295 = 9 * 255, the largest sum the 3x3 area pooling divides), q = RN(x * y), q' = fma(fma(-3, q, x), y, q)
 * with y = RN(1/3) equals the IEEE division x / 3.  argv[1] = stride over the bit patterns (1 = all 1.16e9 of them).
 * Prints the number of mismatches. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char** argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u;
    const float y = 1.0f / 3.0f, top = 2295.0f;
    uint32_t hb;
    memcpy(&hb, &top, 4);
    unsigned long long bad = 0, n = 0;
    for (uint32_t b = 0; b <= hb + 16u; b += stride, n++) {
        float x;
        memcpy(&x, &b, 4);
        const float q = x * y;
        const float q1 = fmaf(fmaf(-3.0f, q, x), y, q);
        if (q1 != x / 3.0f) bad++;
    }
    printf("%llu %llu\n", bad, n);
    return 0;
}
