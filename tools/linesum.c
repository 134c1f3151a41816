/* linesum -- order-independent digest of a text stream: number of lines, and the sum and the xor of a 64-bit hash of every line.
 * Used to compare a full-size `minimod freq` table (hundreds of millions of rows, no defined order among rows sharing
 * (contig,pos)) with the table the unmodified reference printed for the same job, without sorting either.
 * Test infrastructure (tests/test_gpu_fullsize.py, tools/make_fullsize_checksums.sh). */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

int main(void) {
    static char buf[1 << 22];
    uint64_t n = 0, sum = 0, x = 0, h = 1469598103934665603ull;
    size_t got;
    int open_line = 0;
    while ((got = fread(buf, 1, sizeof(buf), stdin)) > 0) {
        for (size_t i = 0; i < got; ++i) {
            const unsigned char c = (unsigned char)buf[i];
            if (c == '\n') {
                h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
                ++n; sum += h; x ^= h; h = 1469598103934665603ull; open_line = 0;
            } else { h = (h ^ c) * 1099511628211ull; open_line = 1; }
        }
    }
    if (open_line) { h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32; ++n; sum += h; x ^= h; }
    printf("%llu %016llx %016llx\n", (unsigned long long)n, (unsigned long long)sum, (unsigned long long)x);
    return 0;
}
