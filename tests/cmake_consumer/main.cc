// Links against the installed package like ggml does; exercises only host-side entry points (no GPU needed).
#include <cstdio>
#include <cstring>
#include "tmac_b200.h"
#include "t-mac/tmac_gemm_wrapper.h"

#define STR2(x) #x
#define STR(x) STR2(x)

int main() {
#ifndef GGML_TMAC_B200
#error "the package must define GGML_TMAC_B200"
#endif
    if (tmac_b200_version() <= 0) return 1;
#ifdef TMAC_KCFG_FILE
    const int n = tmac_b200_load_kcfg_file(STR(TMAC_KCFG_FILE));
    std::printf("kcfg sections: %d\n", n);
    if (n <= 0) return 2;
    tmac_b200_kcfg c;
    if (tmac_b200_find_kcfg(4096 * 2, 4096, 2, &c) != 0 || c.M != 4096) return 3;
#else
#error "TMAC_KCFG_FILE should be defined when kcfg.ini is installed"
#endif
    if (ggml_tmac_get_type_bits(37) != 2) return 4;
    std::printf("consumer ok\n");
    return 0;
}
