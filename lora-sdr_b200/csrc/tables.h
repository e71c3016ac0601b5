// tables.h -- host-side table recipes of the demod hot path (see tables.cpp).
#pragma once
#include <cstddef>
#include <vector>

namespace lorab200 {

struct cfloat { float re, im; };

constexpr size_t kFineSteps = 128; // LoRaDemod.cpp:70  _fineSteps(128)

void build_dechirp_tables(int sf, std::vector<cfloat> &up, std::vector<cfloat> &down);
void build_fine_table(int sf, std::vector<cfloat> &fine);
void build_twiddles(int sf, std::vector<cfloat> &tw);
float power_scale(int sf);
int gen_chirp(cfloat *samps, int N, int ovs, int NN, float f0, bool down, float ampl,
              float &phase_accum);

} // namespace lorab200
