// Host-only harness around csrc/zb200_basis_math.h: evaluates the same per-pixel
// arithmetic the CUDA basis generator runs, so the CPU test-suite can check the
// recurrence against the oracle without a GPU.  Usage: basis_host n_max size out.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "zb200_basis_math.h"

int main(int argc, char** argv) {
    if (argc != 4) return 2;
    const int n_max = atoi(argv[1]), k = atoi(argv[2]);
    const int M = (n_max + 1) * (n_max + 2) / 2;
    std::vector<double> out((size_t)M * k * k, 0.0);
    for (int row = 0; row < k; ++row)
        for (int col = 0; col < k; ++col) {
            const double x = zb200::grid_coord(col, k), y = zb200::grid_coord(row, k);
            const double rho = zb200::grid_rho(x, y);
            if (!(rho <= 1.0)) continue;
            const double theta = atan2(y, x);
            for (int am = 0; am <= n_max; ++am) {
                const double c = cos(am * theta), s = sin(am * theta);
                zb200::RadialIter it(rho, am);
                for (int n = am; n <= n_max; n += 2) {
                    const double r = it.value() * zb200::mode_norm(n, am);
                    out[((size_t)zb200::mode_index(n, am) * k + row) * k + col] = r * c;
                    if (am > 0) out[((size_t)zb200::mode_index(n, -am) * k + row) * k + col] = r * s;
                    it.next();
                }
            }
        }
    FILE* f = fopen(argv[3], "wb");
    if (!f) return 3;
    fwrite(out.data(), sizeof(double), out.size(), f);
    fclose(f);
    return 0;
}
