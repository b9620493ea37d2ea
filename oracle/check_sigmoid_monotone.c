/* check_sigmoid_monotone.c -- TEST INFRASTRUCTURE (oracle/): walks every binary32 value from -104 to 90 in increasing
 * order and checks that y3d_sigmoidf (oracle/y3d_oracle.c; the GPU's dm::sigmoid_ is the same operation sequence) never
 * decreases.  Outside that range the function is constant (0 below, 1 above).  About 100 s on one core:
 *     gcc -O2 -ffp-contract=off -fopenmp -o /tmp/chk oracle/check_sigmoid_monotone.c oracle/y3d_oracle.c -lm && /tmp/chk
 * Result recorded in DESIGN.md: 2 240 020 482 values, 0 decreases. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

float y3d_sigmoidf(float x);

int main(void) {
    long bad = 0, n = 0;
    float prev = -1.0f, lo = -104.0f, hi = 90.0f;
    uint32_t b0, b1;
    memcpy(&b0, &lo, 4);
    memcpy(&b1, &hi, 4);
    for (uint32_t b = b0;; --b) { /* negative floats: the bit pattern decreases as the value increases */
        float x, s;
        memcpy(&x, &b, 4);
        s = y3d_sigmoidf(x);
        if (s < prev) ++bad;
        prev = s;
        ++n;
        if (b == 0x80000000u) break;
    }
    for (uint32_t b = 0; b <= b1; ++b) {
        float x, s;
        memcpy(&x, &b, 4);
        s = y3d_sigmoidf(x);
        if (s < prev) ++bad;
        prev = s;
        ++n;
    }
    printf("values %ld decreases %ld\n", n, bad);
    return bad != 0;
}
