// codec.h -- symbols -> payload bytes: the integer half of the
// receiver, per packet: Gray map, diagonal de-interleave, de-whitening, Hamming / parity
// decode, explicit header, CRC.  Restates LoRaDecoder.cpp:196-397 with the helpers of
// LoRaCodes.hpp (line references at each function).  Written once for host and device: the
// device kernel (decode.cuh) runs one packet per thread; the host build exists only so the
// logic can be checked against the unmodified reference decoder on a machine without a GPU
// (oracle/codec_host.cpp -- test infrastructure, the product never decodes on the CPU).
//
// Everything is bit-exact integer work; the quirks of the reference are kept on purpose
// (e.g. the de-interleaver of the payload uses the CONFIGURED coding rate while the Hamming
// stage uses the rate found in the explicit header, LoRaDecoder.cpp:236,296; with hdr off an
// explicit packet is shortened by 5 bytes whether or not it carries a CRC, :381-384).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LB_HD __host__ __device__ __forceinline__
#else
#define LB_HD inline
#endif

namespace lorab200 {

struct DecoderCfg {
    int sf;            // setSpreadFactor
    int ppm;           // setSymbolSize, 0 = sf
    int rdd;           // setCodingRate: "4/4".."4/8" -> 0..4
    int explicit_hdr;  // enableExplicit
    int hdr;           // enableHdr: keep the 3 header bytes (and the CRC) in the output
    int data_length;   // setDataLength (implicit mode)
    int crcc;          // enableCrcc
    int error_check;   // enableErrorCheck
};

constexpr int kHeaderRdd = 4;        // LoRaCodes.hpp:4   HEADER_RDD
constexpr int kHeaderSymbols = 8;    // LoRaCodes.hpp:5   N_HEADER_SYMBOLS
constexpr int kHeaderCodewords = 5;  // LoRaCodes.hpp:6   N_HEADER_CODEWORDS

// result codes of decode_packet (>= 0: number of output bytes)
constexpr int kDecodeDropped = -1;   // LoRaDecoder::drop()
constexpr int kDecodeIgnored = -2;   // fewer than 8 symbols: the reference returns without output or drop (:208)
constexpr int kDecodeNoRoom = -3;    // output buffer too small (ours)

LB_HD unsigned codec_round_up(unsigned num, unsigned factor) { return ((num + factor - 1) / factor) * factor; } // LoRaCodes.hpp:12-15

// scratch a packet of n symbols needs: codewords, then bytes
LB_HD unsigned codec_num_symbols(const DecoderCfg &c, unsigned n_in) { return codec_round_up(n_in, 4u + unsigned(c.rdd)); }
LB_HD unsigned codec_num_codewords(const DecoderCfg &c, unsigned n_in)
{
    const unsigned ppm = unsigned(c.ppm ? c.ppm : c.sf);
    return codec_num_symbols(c, n_in) / (4u + unsigned(c.rdd)) * ppm;                    // LoRaDecoder.cpp:211
}

// LoRaCodes.hpp:29-53
LB_HD uint8_t codec_header_checksum(const uint8_t *h)
{
    const unsigned a = h[0] >> 4, b = h[0] & 0xf, c = h[1] & 0xf;
    auto bit = [](unsigned v, int i) { return (v >> i) & 1u; };
    unsigned r = (bit(a, 0) ^ bit(a, 1) ^ bit(a, 2) ^ bit(a, 3)) << 4;
    r |= (bit(a, 3) ^ bit(b, 1) ^ bit(b, 2) ^ bit(b, 3) ^ bit(c, 0)) << 3;
    r |= (bit(a, 2) ^ bit(b, 0) ^ bit(b, 3) ^ bit(c, 1) ^ bit(c, 3)) << 2;
    r |= (bit(a, 1) ^ bit(b, 0) ^ bit(b, 2) ^ bit(c, 0) ^ bit(c, 1) ^ bit(c, 2)) << 1;
    r |= bit(a, 0) ^ bit(b, 1) ^ bit(c, 0) ^ bit(c, 1) ^ bit(c, 2) ^ bit(c, 3);
    return uint8_t(r);
}

// LoRaCodes.hpp:55-93: CCITT-style CRC masked with an 8-bit LFSR
LB_HD uint16_t codec_data_checksum(const uint8_t *data, int length)
{
    auto crc_step = [](uint16_t crc) {
        for (int i = 0; i < 8; i++) crc = (crc & 0x8000) ? uint16_t((crc << 1) ^ 0x1021) : uint16_t(crc << 1);
        return crc;
    };
    auto xsum = [](uint8_t t) { t ^= t >> 4; t ^= t >> 2; t ^= t >> 1; return uint8_t(t & 1); };
    uint16_t res = 0;
    uint8_t v = 0xff;
    for (int i = 0; i < length; i++) {
        const uint16_t crc = crc_step(res);
        v = uint8_t(xsum(v & 0xB8) | (v << 1));
        res = uint16_t(crc ^ data[i]);
    }
    res ^= v;
    v = uint8_t(xsum(v & 0xB8) | (v << 1));
    res ^= uint16_t(v << 8);
    return res;
}

// LoRaCodes.hpp:152-166: the interleaved whitening LFSRs (poly 0x1D), count is a uint16 in the reference
LB_HD void codec_dewhiten(uint8_t *buf, unsigned count, int bit_ofs, int rdd)
{
    count &= 0xffffu;
    const uint8_t m = uint8_t(0xff >> (4 - rdd));
    uint64_t r[2];
    r[0] = (rdd == 1) ? 0x05121100F8ECFEEFull : 0x6572D100E85C2EFFull;
    r[1] = (rdd == 1) ? 0xF8ECFEEFEFEFEFEFull : 0xE85C2EFFFFFFFFFFull;
    auto step = [](uint64_t x) { return (x >> 8) | (((x >> 32) ^ (x >> 24) ^ (x >> 16) ^ x) << 56); };
    int i = 0;
    for (; i < bit_ofs; i++) r[i & 1] = step(r[i & 1]);
    for (unsigned j = 0; j < count; j++, i++) {
        buf[j] ^= uint8_t(r[i & 1]) & m;
        r[i & 1] = step(r[i & 1]);
    }
}

// LoRaCodes.hpp:210-242 (8,4) and :267-296 (7,4): single-error correcting, sx1272 bit order
LB_HD uint8_t codec_hamming84(uint8_t b, bool &error, bool &bad)
{
    auto bit = [b](int i) { return unsigned(b >> i) & 1u; };
    const unsigned parity = (bit(0) ^ bit(1) ^ bit(2) ^ bit(4)) | (bit(1) ^ bit(2) ^ bit(3) ^ bit(5)) << 1 |
                            (bit(0) ^ bit(1) ^ bit(3) ^ bit(6)) << 2 | (bit(0) ^ bit(2) ^ bit(3) ^ bit(7)) << 3;
    if (parity) error = true;
    switch (parity) {
    case 0xD: return (b ^ 1) & 0xf;
    case 0x7: return (b ^ 2) & 0xf;
    case 0xB: return (b ^ 4) & 0xf;
    case 0xE: return (b ^ 8) & 0xf;
    case 0x0: case 0x1: case 0x2: case 0x4: case 0x8: return b & 0xf;
    default: bad = true; return b & 0xf;
    }
}
LB_HD uint8_t codec_hamming74(uint8_t b, bool &error)
{
    auto bit = [b](int i) { return unsigned(b >> i) & 1u; };
    const unsigned parity = (bit(0) ^ bit(1) ^ bit(2) ^ bit(4)) | (bit(1) ^ bit(2) ^ bit(3) ^ bit(5)) << 1 |
                            (bit(0) ^ bit(1) ^ bit(3) ^ bit(6)) << 2;
    if (parity) error = true;
    switch (parity) {
    case 0x5: return (b ^ 1) & 0xf;
    case 0x7: return (b ^ 2) & 0xf;
    case 0x3: return (b ^ 4) & 0xf;
    case 0x6: return (b ^ 8) & 0xf;
    default: return b & 0xf;
    }
}
// LoRaCodes.hpp:302-307 and :319-327: parity checks of the 4/5 and 4/6 codes
LB_HD uint8_t codec_parity54(uint8_t b, bool &error)
{
    unsigned x = b ^ (b >> 2);
    x = x ^ (x >> 1) ^ (b >> 4);
    if (x & 1) error = true;
    return b & 0xf;
}
LB_HD uint8_t codec_parity64(uint8_t b, bool &error)
{
    unsigned x = b ^ (b >> 1) ^ (b >> 2);
    unsigned y = x ^ b ^ (b >> 3);
    x ^= b >> 4;
    y ^= b >> 5;
    if ((x | y) & 1) error = true;
    return b & 0xf;
}

// one codeword -> nibble under coding rate rdd (the five branches of LoRaDecoder.cpp:325-361)
LB_HD uint8_t codec_nibble(uint8_t cw, int rdd, bool &error, bool &bad)
{
    switch (rdd) {
    case 0: return cw;                       // (the callers mask or shift as the reference does)
    case 1: return codec_parity54(cw, error);
    case 2: return codec_parity64(cw, error);
    case 3: return codec_hamming74(cw, error) & 0xf;
    default: return codec_hamming84(cw, error, bad) & 0xf;
    }
}

// LoRaCodes.hpp:357-374: symbol k of a block carries bit k of every codeword, rotated by k
LB_HD void codec_deinterleave(const uint16_t *sym, unsigned n_sym, uint8_t *cw, unsigned ppm, unsigned rdd)
{
    const unsigned per = 4 + rdd;
    for (unsigned x = 0; x < n_sym / per; x++)
        for (unsigned k = 0; k < per; k++) {
            const unsigned s = sym[x * per + k];
            for (unsigned m = 0; m < ppm; m++) cw[x * ppm + (m + k) % ppm] |= uint8_t(((s >> m) & 1u) << k);
        }
}

// Byte scratch of decode_packet for n_cw codewords: one byte more than the n_cw / 2 the payload
// can occupy.  With an explicit header, an even PPM and a packet of exactly one block the (8,4)
// nibbles of the first block end on byte n_cw / 2 (LoRaDecoder.cpp:315-340 writes one byte past its
// vector there); that byte is never part of the output, but it must not be another packet's.
LB_HD unsigned codec_scratch_bytes(unsigned n_cw) { return n_cw / 2 + 1; }

// LoRaDecoder.cpp:196-397.  `sym_io` holds n_in received symbols followed by room up to
// codec_num_symbols() entries (it is modified in place), `cw` codec_num_codewords() bytes and
// `bytes` codec_scratch_bytes(codec_num_codewords()) bytes of scratch.  Returns the payload length
// written to `out`, or one of the kDecode* codes.
LB_HD int decode_packet(const DecoderCfg &c, uint16_t *sym_io, unsigned n_in, uint8_t *cw, uint8_t *bytes,
                        uint8_t *out, unsigned out_cap)
{
    const unsigned sf = unsigned(c.sf), ppm = unsigned(c.ppm ? c.ppm : c.sf);
    if (n_in < unsigned(kHeaderSymbols)) return kDecodeIgnored;                          // :208
    const unsigned n_sym = codec_num_symbols(c, n_in), n_cw = codec_num_codewords(c, n_in);
    const unsigned n_bytes = (n_cw + 1) / 2;
    int rdd = c.rdd;

    // Gray map after rounding the (sf - ppm) low bits away (:218-222); unsigned-short arithmetic
    for (unsigned i = 0; i < n_sym; i++) {
        uint16_t s = (i < n_in) ? sym_io[i] : uint16_t(0);
        s = uint16_t(s + (1u << (sf - ppm)) / 2);
        s = uint16_t(s >> (sf - ppm));
        sym_io[i] = uint16_t(s ^ (s >> 1));
    }
    for (unsigned i = 0; i < n_cw; i++) cw[i] = 0;
    for (unsigned i = 0; i < n_bytes; i++) bytes[i] = 0;

    // de-interleave + de-whiten (:224-259): the first 8 symbols always use the header rate
    if (rdd != kHeaderRdd) {
        codec_deinterleave(sym_io, kHeaderSymbols, cw, ppm, kHeaderRdd);
        if (c.explicit_hdr) codec_dewhiten(cw + kHeaderCodewords, ppm - kHeaderCodewords, 0, kHeaderRdd);
        else codec_dewhiten(cw, ppm, 0, kHeaderRdd);
        if (n_sym > unsigned(kHeaderSymbols)) {
            codec_deinterleave(sym_io + kHeaderSymbols, n_sym - kHeaderSymbols, cw + ppm, ppm, unsigned(rdd));
            codec_dewhiten(cw + ppm, n_cw - ppm, int(c.explicit_hdr ? ppm - kHeaderCodewords : ppm), rdd);
        }
    } else {
        codec_deinterleave(sym_io, n_sym, cw, ppm, unsigned(rdd));
        if (c.explicit_hdr) codec_dewhiten(cw + kHeaderCodewords, n_cw - kHeaderCodewords, 0, rdd);
        else codec_dewhiten(cw, n_cw, 0, rdd);
    }

    bool error = false, bad = false;
    unsigned d = 0, k = 0;                 // nibble offset into bytes, codeword offset
    unsigned packet_len = 0, data_len = 0;
    bool check_crc = c.crcc != 0;

    if (c.explicit_hdr) {                  // :279-303
        bytes[0] = uint8_t(codec_hamming84(cw[1], error, bad) & 0xf);
        bytes[0] |= uint8_t(codec_hamming84(cw[0], error, bad) << 4);
        bytes[1] = uint8_t(codec_hamming84(cw[2], error, bad) & 0xf);
        bytes[2] = uint8_t(codec_hamming84(cw[4], error, bad) & 0xf);
        bytes[2] |= uint8_t(codec_hamming84(cw[3], error, bad) << 4);
        bytes[2] ^= codec_header_checksum(bytes);
        if (error && c.error_check) return kDecodeDropped;
        if ((bytes[1] & 1) == 0) check_crc = false;
        rdd = (bytes[1] >> 1) & 0x7;
        if (rdd > 4) return kDecodeDropped;
        packet_len = bytes[0];
        data_len = packet_len + ((bytes[1] & 1) ? 5u : 3u);
        k = kHeaderCodewords;
        d = 6;
    } else {                               // :304-311
        packet_len = unsigned(c.data_length);
        data_len = c.crcc ? packet_len + 2 : packet_len;
    }
    if (data_len > n_bytes) return kDecodeDropped;                                       // :313

    // the rest of the first block is always (8,4) coded (:315-320)
    for (; k < ppm; k++, d++) {
        const uint8_t nib = codec_hamming84(cw[k], error, bad);
        if (d & 1) bytes[d >> 1] |= uint8_t(nib << 4);
        else bytes[d >> 1] = uint8_t(nib & 0xf);
    }
    // an odd nibble count is completed with the first codeword of the payload rate (:322-340)
    if (d & 1) {
        const uint8_t nib = (k < n_cw) ? codec_nibble(cw[k], rdd, error, bad) : uint8_t(0);
        k++;
        bytes[d >> 1] |= uint8_t(nib << 4);
        d++;
    }
    d >>= 1;
    if (error && c.error_check) return kDecodeDropped;                                   // :342

    // payload: two codewords per byte (:345-361)
    for (unsigned i = d; i < data_len; i++) {
        const uint8_t lo = (k < n_cw) ? codec_nibble(cw[k], rdd, error, bad) : uint8_t(0);
        const uint8_t hi = (k + 1 < n_cw) ? codec_nibble(cw[k + 1], rdd, error, bad) : uint8_t(0);
        k += 2;
        bytes[i] = (rdd == 0) ? uint8_t((lo & 0xf) | uint8_t(hi << 4)) : uint8_t(lo | uint8_t(hi << 4));
    }
    if (error && c.error_check) return kDecodeDropped;                                   // :363

    unsigned first = 0;
    if (c.explicit_hdr) {                  // :367-380
        if (bytes[1] & 1) {
            const uint16_t crc = codec_data_checksum(bytes + 3, int(packet_len));
            const uint16_t got = uint16_t(bytes[3 + packet_len] | (bytes[4 + packet_len] << 8));
            if (crc != got && check_crc) return kDecodeDropped;
            bytes[3 + packet_len] ^= uint8_t(crc);
            bytes[4 + packet_len] ^= uint8_t(crc >> 8);
        }
        if (!c.hdr) { first = 3; data_len -= 5; }
    } else if (check_crc) {                // :381-389
        const uint16_t crc = codec_data_checksum(bytes, c.data_length);
        const uint16_t got = uint16_t(bytes[c.data_length] | (bytes[c.data_length + 1] << 8));
        if (crc != got) return kDecodeDropped;
        bytes[c.data_length] ^= uint8_t(crc);
        bytes[c.data_length + 1] ^= uint8_t(crc >> 8);
    }
    if (data_len > out_cap) return kDecodeNoRoom;
    for (unsigned i = 0; i < data_len; i++) out[i] = bytes[first + i];
    return int(data_len);
}

// ---- encoder (LoRaEncoder.cpp:137-228): payload bytes -> symbols ----------------------------
struct EncoderCfg {
    int sf;            // setSpreadFactor
    int ppm;           // setSymbolSize, 0 = sf
    int rdd;           // setCodingRate
    int explicit_hdr;  // enableExplicit
    int crc;           // enableCrc
    int whitening;     // enableWhitening
};

LB_HD unsigned codec_enc_codewords(const EncoderCfg &c, unsigned payload_len)
{
    const unsigned ppm = unsigned(c.ppm ? c.ppm : c.sf);
    return codec_round_up((payload_len + (c.crc ? 2u : 0u)) * 2u + (c.explicit_hdr ? unsigned(kHeaderCodewords) : 0u), ppm); // :186
}
LB_HD unsigned codec_enc_symbols(const EncoderCfg &c, unsigned payload_len)
{
    const unsigned ppm = unsigned(c.ppm ? c.ppm : c.sf);
    return unsigned(kHeaderSymbols) + (codec_enc_codewords(c, payload_len) / ppm - 1u) * (4u + unsigned(c.rdd));              // :187
}

// LoRaCodes.hpp:190-204 / :250-262 / :309-313 / :329-333: the encoders of the four codes
LB_HD uint8_t codec_encode_nibble(uint8_t x, int rdd)
{
    const unsigned d0 = x & 1u, d1 = (x >> 1) & 1u, d2 = (x >> 2) & 1u, d3 = (x >> 3) & 1u;
    unsigned b = x & 0xfu;
    switch (rdd) {
    case 0: return uint8_t(x);                                   // (the caller passes the nibble)
    case 1: { unsigned t = x ^ (x >> 2); t = t ^ (t >> 1); return uint8_t((x & 0xf) | ((t << 4) & 0x10)); }
    case 2: { const unsigned u = x ^ (x >> 1) ^ (x >> 2); const unsigned v = u ^ x ^ (x >> 3);
              return uint8_t(((u & 1) << 4) | ((v & 1) << 5) | (x & 0xf)); }
    case 3: b |= (d0 ^ d1 ^ d2) << 4; b |= (d1 ^ d2 ^ d3) << 5; b |= (d0 ^ d1 ^ d3) << 6; return uint8_t(b);
    default: b |= (d0 ^ d1 ^ d2) << 4; b |= (d1 ^ d2 ^ d3) << 5; b |= (d0 ^ d1 ^ d3) << 6; b |= (d0 ^ d2 ^ d3) << 7; return uint8_t(b);
    }
}

// LoRaCodes.hpp:336-351: codeword bit k of a block goes to symbol k, rotated by k
LB_HD void codec_interleave(const uint8_t *cw, unsigned n_cw, uint16_t *sym, unsigned ppm, unsigned rdd)
{
    const unsigned per = 4 + rdd;
    for (unsigned x = 0; x < n_cw / ppm; x++)
        for (unsigned k = 0; k < per; k++) {
            unsigned s = sym[x * per + k];
            for (unsigned m = 0; m < ppm; m++) s |= ((unsigned(cw[x * ppm + (m + k) % ppm]) >> k) & 1u) << m;
            sym[x * per + k] = uint16_t(s);
        }
}

// LoRaEncoder.cpp:137-228.  `bytes` : payload_len + 2 bytes of scratch, `cw`: codec_enc_codewords()
// bytes of scratch, `sym`: codec_enc_symbols() outputs.  Nibbles past the end of the payload (the
// padding that rounds the codewords up to a multiple of PPM) are ZERO here; the reference reads them
// from beyond its byte vector (LoRaEncoder.cpp:141-146 with count > what is left), i.e. they are
// unspecified there.  Returns the number of symbols.
LB_HD int encode_packet(const EncoderCfg &c, const uint8_t *payload, unsigned payload_len, uint8_t *bytes,
                        uint8_t *cw, uint16_t *sym)
{
    const unsigned sf = unsigned(c.sf), ppm = unsigned(c.ppm ? c.ppm : c.sf);
    const unsigned total = payload_len + (c.crc ? 2u : 0u);
    const unsigned n_cw = codec_enc_codewords(c, payload_len), n_sym = codec_enc_symbols(c, payload_len);
    for (unsigned i = 0; i < payload_len; i++) bytes[i] = payload[i];
    if (c.crc) {
        const uint16_t crc = codec_data_checksum(bytes, int(payload_len));
        bytes[payload_len] = uint8_t(crc);
        bytes[payload_len + 1] = uint8_t(crc >> 8);
    }
    for (unsigned i = 0; i < n_cw; i++) cw[i] = 0;
    for (unsigned i = 0; i < n_sym; i++) sym[i] = 0;
    unsigned k = 0, d = 0;
    if (c.explicit_hdr) {                                                   // :198-211
        uint8_t h[3];
        h[0] = uint8_t(payload_len);
        h[1] = uint8_t((c.crc ? 1 : 0) | (c.rdd << 1));
        h[2] = codec_header_checksum(h);
        cw[k++] = codec_encode_nibble(h[0] >> 4, 4);
        cw[k++] = codec_encode_nibble(h[0] & 0xf, 4);
        cw[k++] = codec_encode_nibble(h[1] & 0xf, 4);
        cw[k++] = codec_encode_nibble(h[2] >> 4, 4);
        cw[k++] = codec_encode_nibble(h[2] & 0xf, 4);
    }
    auto nibble = [&](unsigned dd) -> uint8_t {
        if ((dd >> 1) >= total) return 0;                                   // padding (see above)
        return (dd & 1) ? uint8_t(bytes[dd >> 1] >> 4) : uint8_t(bytes[dd >> 1] & 0xf);
    };
    const unsigned k1 = k;
    for (; k < ppm; k++, d++) cw[k] = codec_encode_nibble(nibble(d), 4);    // :213  first block: always (8,4)
    if (c.whitening) codec_dewhiten(cw + k1, ppm - k1, 0, kHeaderRdd);      // :214-216
    if (n_cw > ppm) {                                                       // :218-224
        for (; k < n_cw; k++, d++) cw[k] = codec_encode_nibble(nibble(d), c.rdd);
        if (c.whitening) codec_dewhiten(cw + ppm, n_cw - ppm, int(ppm - k1), c.rdd);
    }
    codec_interleave(cw, ppm, sym, ppm, kHeaderRdd);                        // :227-231
    if (n_cw > ppm) codec_interleave(cw + ppm, n_cw - ppm, sym + kHeaderSymbols, ppm, unsigned(c.rdd));
    for (unsigned i = 0; i < n_sym; i++) {                                  // :233-237  Gray -> binary, pad the LSBs
        uint16_t g = sym[i];
        g = uint16_t(g ^ (g >> 8)); g = uint16_t(g ^ (g >> 4)); g = uint16_t(g ^ (g >> 2)); g = uint16_t(g ^ (g >> 1));
        sym[i] = uint16_t(g << (sf - ppm));
    }
    return int(n_sym);
}

} // namespace lorab200
