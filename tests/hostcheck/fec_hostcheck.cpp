// Host build of p25rx_b200/csrc/p25_fec.cuh -- TEST HARNESS ONLY.
// Lets the CPU-only test suite exercise the exact decoder source the GPU compiles, against the
// oracle, on machines without a GPU.  It is never loaded by the p25rx_b200 package and is not a
// fallback: the product library has no host compute path.
#define P25_FEC_HOSTCHECK 1
#include "../../p25rx_b200/csrc/p25_fec.cuh"

static P25DevTables g_T;
static bool g_init = false;
static const P25DevTables& T() {
    if (!g_init) {
        p25_fill_tables(&g_T);
        g_init = true;
    }
    return g_T;
}

extern "C" {
int hc_bch_decode(uint64_t w, uint32_t* d) { return p25_bch_decode(T(), w, d); }
int hc_golay23_decode(uint32_t w, uint32_t* d) { return p25_golay23_decode(T(), w, d); }
int hc_golay24_decode(uint32_t w, uint32_t* d) { return p25_golay24_decode(T(), w, d); }
int hc_golay18_decode(uint32_t w, uint32_t* d) { return p25_golay18_decode(T(), w, d); }
int hc_hamming15_decode(uint32_t w, uint32_t* d) { return p25_hamming15_decode(T(), w, d); }
int hc_hamming10_decode(uint32_t w, uint32_t* d) { return p25_hamming10_decode(T(), w, d); }
int hc_cyclic16_decode(uint32_t w, uint32_t* d) { return p25_cyclic16_decode(T(), w, d); }
int hc_rs_decode(uint8_t* sym, int n, int k) { return p25_rs_decode(T(), sym, n, k); }
int hc_trellis_half_decode(const uint8_t* d98, uint8_t* out12) { return p25_trellis_half_decode(T(), d98, out12); }
int hc_trellis_34_decode(const uint8_t* d98, uint8_t* out18) { return p25_trellis_34_decode(T(), d98, out18); }
uint32_t hc_crc_ccitt(const uint8_t* d, int n) { return p25_crc_ccitt(d, n); }
void hc_imbe_decode(const uint8_t* d72, uint32_t* c, uint32_t* e) { p25_imbe_decode(T(), d72, c, e); }
}
