// tcgen05 tensor-core decoder engine (placeholder until the kernels land).
#pragma once
#include <string>
#include "dspgn_simt.cuh"

namespace dspgn {
constexpr int kTcRows = 128;
struct TcDecoderHost { bool ok = false; void* blob = nullptr; };
inline int tc_pack_decoder(const DspgnDecoderSpec&, const float* const*, const float* const*, TcDecoderHost& h,
                           DecoderDev* dv, std::string&) { h.ok = false; dv->tc_blob = nullptr; return 0; }
inline void tc_free_decoder(TcDecoderHost&) {}
inline int tc_setup_kernels(std::string&) { return 0; }
inline bool tc_engine_default() { return false; }
inline int tc_launch_term(TermArgs&, int, long long, cudaStream_t, std::string& err) { err = "tc engine not built"; return DSPGN_E_ARG; }
}  // namespace dspgn
