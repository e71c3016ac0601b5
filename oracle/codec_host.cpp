/* codec_host.cpp -- TEST INFRASTRUCTURE ONLY.  Builds the product's symbol->byte codec
 * (lora-sdr_b200/csrc/codec.h, a host+device header) for the HOST, so that its logic can be
 * compared with the unmodified reference decoder on a machine without a GPU.  The product
 * itself only decodes on the device (lorab200_decode_dev). */
#include <cstdint>
#include <cstring>
#include <vector>

#include "../lora-sdr_b200/csrc/codec.h"

extern "C" int codec_host_decode(const int32_t *cfg8 /* sf, ppm, rdd, explicit, hdr, data_length, crcc, error_check */,
                                 const uint16_t *symbols, int n_symbols, uint8_t *out, int out_cap)
{
    lorab200::DecoderCfg c;
    c.sf = cfg8[0]; c.ppm = cfg8[1]; c.rdd = cfg8[2]; c.explicit_hdr = cfg8[3];
    c.hdr = cfg8[4]; c.data_length = cfg8[5]; c.crcc = cfg8[6]; c.error_check = cfg8[7];
    const unsigned n = unsigned(n_symbols < 0 ? 0 : n_symbols);
    std::vector<uint16_t> sym(lorab200::codec_num_symbols(c, n) + 16);
    std::memcpy(sym.data(), symbols, size_t(n) * 2);
    // exact sizes plus one guard byte each: an overrun of either scratch is reported as -100
    const unsigned n_cw = lorab200::codec_num_codewords(c, n), n_b = lorab200::codec_scratch_bytes(n_cw);
    std::vector<uint8_t> cw(n_cw + 1), bytes(n_b + 1);
    cw[n_cw] = 0xA5; bytes[n_b] = 0x5A;
    const int r = lorab200::decode_packet(c, sym.data(), n, cw.data(), bytes.data(), out, unsigned(out_cap));
    if (cw[n_cw] != 0xA5 || bytes[n_b] != 0x5A) return -100;
    return r;
}

extern "C" int codec_host_encode(const int32_t *cfg6 /* sf, ppm, rdd, explicit, crc, whitening */,
                                 const uint8_t *payload, int length, uint16_t *symbols, int max_symbols)
{
    lorab200::EncoderCfg c;
    c.sf = cfg6[0]; c.ppm = cfg6[1]; c.rdd = cfg6[2]; c.explicit_hdr = cfg6[3]; c.crc = cfg6[4]; c.whitening = cfg6[5];
    const unsigned n = unsigned(length);
    if (int(lorab200::codec_enc_symbols(c, n)) > max_symbols) return -1;
    std::vector<uint8_t> bytes(n + 2), cw(lorab200::codec_enc_codewords(c, n));
    return lorab200::encode_packet(c, payload, n, bytes.data(), cw.data(), symbols);
}

/* ---- the receiver state machine (lora-sdr_b200/csrc/rx_fsm.h), host build, for tests ---------- */
#include "../lora-sdr_b200/csrc/rx_fsm.h"

extern "C" int rx_host_needs_second(const int32_t *cfg4 /* n, sync, mtu, thresh as float bits */, const void *state,
                                    const lorab200_result *r0)
{
    lorab200::RxCfg c;
    c.n = cfg4[0]; c.sync = cfg4[1]; c.mtu = cfg4[2];
    std::memcpy(&c.thresh, &cfg4[3], 4);
    return lorab200::rx_needs_second(c, *static_cast<const lorab200::RxState *>(state), *r0) ? 1 : 0;
}

/* returns consume | emit_symbol | post_packet through out3 */
extern "C" void rx_host_step(const int32_t *cfg4, void *state, const lorab200_result *r0, int idx_after0, int have_r1,
                             const lorab200_result *r1, int64_t *out3)
{
    lorab200::RxCfg c;
    c.n = cfg4[0]; c.sync = cfg4[1]; c.mtu = cfg4[2];
    std::memcpy(&c.thresh, &cfg4[3], 4);
    const lorab200::RxAction a = lorab200::rx_step(c, *static_cast<lorab200::RxState *>(state), *r0, idx_after0,
                                                   have_r1 != 0, *r1);
    out3[0] = a.consume; out3[1] = a.emit_symbol; out3[2] = a.post_packet;
}

/* ---- the fine-tune index orbit (lora-sdr_b200/csrc/fine_orbit.h), host build, for tests ---------- */
#include "../lora-sdr_b200/csrc/fine_orbit.h"
#include "../lora-sdr_b200/csrc/fine_walk.cpp"

/* builds the orbit of (len, d, x0); returns n_seg (0 = no orbit) and evaluates it at the n_q sample numbers q[] */
extern "C" int fine_orbit_host(int32_t len, float d, int32_t x0, const uint64_t *q, int n_q, int32_t *out, int64_t *info3)
{
    const lorab200::FineWalk fw(len, d);
    lorab200::FineOrbit o;
    if (!lorab200::build_fine_orbit(fw, x0, o)) return 0;
    for (int i = 0; i < n_q; ++i) out[i] = lorab200::orbit_at(o, q[i]);
    if (info3) { info3[0] = o.n_pre; info3[1] = o.pre_len; info3[2] = o.period; }
    return o.n_seg;
}
