/* Test harness (not product code): feeds length-prefixed byte strings from a file to the three untrusted-input parsers of
 * liborbx (orbx_wire_parse_frame, orbx_wire_parse_features, orbx_pnm_header), built with -fsanitize=address,undefined by
 * tests/test_wire_fuzz_cpu.py.  Every input is copied into an exactly-sized heap block so that any read past the end trips ASan;
 * pointers returned by a successful parse must lie inside the block.  Prints the three acceptance counts. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "orbx_wire.h"

static int inside(const uint8_t *base, size_t n, const void *p, size_t len) {
    const uint8_t *q = (const uint8_t *)p;
    return q >= base && q <= base + n && len <= (size_t)(base + n - q);
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    long ok_frame = 0, ok_feat = 0, ok_pnm = 0, total = 0;
    for (;;) {
        uint32_t n;
        if (fread(&n, 4, 1, f) != 1) break;
        uint8_t *buf = (uint8_t *)malloc(n ? n : 1);
        if (n && fread(buf, 1, n, f) != n) return 3;
        uint8_t *exact = (uint8_t *)malloc(n ? n : 1);      /* n == 0: a valid pointer, zero readable bytes by contract */
        memcpy(exact, buf, n);
        orbx_wire_frame fr;
        if (orbx_wire_parse_frame(exact, n, &fr) == ORBX_OK) {
            ok_frame++;
            if (!inside(exact, n, fr.type, fr.type_len) || (fr.image && !inside(exact, n, fr.image, fr.image_bytes))) return 10;
        }
        orbx_wire_features ft;
        if (orbx_wire_parse_features(exact, n, &ft) == ORBX_OK) {
            ok_feat++;
            if (ft.n < 0 || !inside(exact, n, ft.keypoints, (size_t)ft.n * sizeof(orbx_keypoint)) || !inside(exact, n, ft.descriptors, (size_t)ft.n * 32)) return 11;
        }
        int w, h, ch; size_t off;
        if (orbx_pnm_header(exact, n, &w, &h, &ch, &off) == ORBX_OK) {
            ok_pnm++;
            if (w <= 0 || h <= 0 || (ch != 1 && ch != 3) || off > n || (unsigned long long)w * h * ch > n - off) return 12;
        }
        free(exact); free(buf);
        total++;
    }
    printf("%ld %ld %ld %ld\n", total, ok_frame, ok_feat, ok_pnm);
    return 0;
}
