// Test harness only: hands the host-side tables of the integer-tensor-pipe decimators (p25rx_b200/csrc/p25_imma_tables.h,
// the header ddc_fm.cu builds its uploads from) to the CPU suite.  Never loaded by the package.
#include <string.h>

#include "../../p25rx_b200/csrc/p25_imma_tables.h"

extern "C" {

// out: [4][18][32][2] words, init[3]; returns ok
int hc_imma5(unsigned* out, int* init, double* gain) {
    static const p25imma::Tables5 t(P25_TAPS_DECIM_H);
    memcpy(out, t.b, sizeof(t.b));
    memcpy(init, t.init, sizeof(t.init));
    *gain = t.gain;
    return t.ok ? 1 : 0;
}

// out: [2][84][32][2] words
int hc_imma50(unsigned* out, int* init, double* gain) {
    static const p25imma::Tables50 t(P25_TAPS_FRONT_H, P25_TAPS_DECIM_H);
    memcpy(out, t.b, sizeof(t.b));
    memcpy(init, t.init, sizeof(t.init));
    *gain = t.gain;
    return t.ok ? 1 : 0;
}

}
