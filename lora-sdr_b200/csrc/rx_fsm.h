// rx_fsm.h -- the reference demod's frame-sync state machine (LoRaDemod.cpp:173-327) as a
// plain state-transition function, host and device (sync
// acquisition on the device).  The receiver kernel (rx_kernel.cuh) runs one instance per
// channel: it demodulates the window at the current stream offset, calls rx_needs_second /
// rx_step, advances the offset, and repeats -- exactly the sequence of work() calls the
// reference block makes, without a host round trip per window.
//
// State starts from zero (the reference reads _fineTuneIndex, _finefreqError and _prevValue
// uninitialised; same convention as the block shell).
#pragma once
#include <stdint.h>

#include "../../include/lorab200.h"

#if defined(__CUDACC__)
#define LB_RX_HD __host__ __device__ __forceinline__
#else
#define LB_RX_HD inline
#endif

namespace lorab200 {

enum RxPhase { RX_FRAMESYNC = 0, RX_DOWNCHIRP0, RX_DOWNCHIRP1, RX_QUARTERCHIRP, RX_DATASYMBOLS };

struct RxCfg {
    int n;            // N = 2^SF
    int sync;         // setSync            LoRaDemod.cpp:124
    float thresh;     // setThreshold (dB)  LoRaDemod.cpp:129
    int mtu;          // setMTU             LoRaDemod.cpp:134
};

// LoRaDemod.cpp:362-392 (the private members that survive a work() call)
struct RxState {
    int32_t phase;          // _state
    int32_t table;          // _chirpTable: LORAB200_TABLE_UP / _DOWN
    int32_t sym_count;      // _symCount
    int32_t prev_value;     // _prevValue (short)
    int32_t freq_error;     // _freqError
    int32_t fine_index;     // _fineTuneIndex
    float   ffe;            // _finefreqError
    // signals emitted at DOWNCHIRP1 for the packet being received (LoRaDemod.cpp:267-269)
    int32_t sig_error;
    float   sig_power, sig_snr;
    // raw detector outputs of that DOWNCHIRP1 window (set by the receiver kernels, not by rx_step):
    // a host can redo the dB conversions with its own libm and get the reference's bits
    float   sig_max_mag2, sig_reserved;
    double  sig_total;
};

// what one transition asks the caller to do
struct RxAction {
    uint32_t consume;       // samples to consume (inPort->consume(total))
    int32_t  emit_symbol;   // >= 0: append this value to the packet under construction
    int32_t  post_packet;   // != 0: the packet is complete (after emit_symbol), post sym_count symbols
};

LB_RX_HD void rx_reset(RxState &s)
{
    s.phase = RX_FRAMESYNC; s.table = LORAB200_TABLE_UP; s.sym_count = 0; s.prev_value = 0; s.freq_error = 0;
    s.fine_index = 0; s.ffe = 0.0f; s.sig_error = 0; s.sig_power = 0.0f; s.sig_snr = 0.0f;
    s.sig_max_mag2 = 0.0f; s.sig_reserved = 0.0f; s.sig_total = 0.0;
}

// LoRaDemod.cpp:181-189: in FRAMESYNC a window that looks like sync word 0 needs the next
// window demodulated too (same table, same fine step, index continuing on a copy)
LB_RX_HD bool rx_needs_second(const RxCfg &c, const RxState &s, const lorab200_result &r0)
{
    if (s.phase != RX_FRAMESYNC) return false;
    const float snr = r0.power - r0.power_avg;
    const bool squelched = snr < c.thresh;
    const bool syncd = !squelched && (s.prev_value + 4) / 8 == 0;
    const bool match0 = (unsigned(r0.index) + 4u) / 8u == unsigned(c.sync >> 4);
    return syncd && match0;
}

// One work() call of the reference.  r0: detector outputs of the window at the current offset;
// idx_after0: _fineTuneIndex after that window's N samples; r1: outputs of the look-ahead
// window when rx_needs_second() said so.
LB_RX_HD RxAction rx_step(const RxCfg &c, RxState &s, const lorab200_result &r0, int idx_after0,
                          bool have_r1, const lorab200_result &r1)
{
    const int N = c.n;
    const unsigned value = unsigned(r0.index);
    float power = r0.power;
    float f_index = r0.f_index;
    const float snr = r0.power - r0.power_avg;
    const bool squelched = snr < c.thresh;                               // :173-174
    s.fine_index = idx_after0;                                            // :160-162 over the window
    RxAction a;
    a.consume = 0; a.emit_symbol = -1; a.post_packet = 0;

    switch (s.phase) {
    case RX_FRAMESYNC: {                                                  // :179-237
        const bool syncd = !squelched && (s.prev_value + 4) / 8 == 0;
        const bool match0 = (value + 4u) / 8u == unsigned(c.sync >> 4);
        bool match1 = false;
        if (syncd && match0 && have_r1) {
            power = r1.power;                                             // detect() overwrites power, powerAvg, fIndex (:203)
            f_index = r1.f_index;
            match1 = (unsigned(r1.index) + 4u) / 8u == unsigned(c.sync & 0xf);
        }
        if (syncd && match0 && match1) {
            a.consume = uint32_t(2 * N);
            s.phase = RX_DOWNCHIRP0;
            s.table = LORAB200_TABLE_DOWN;
        } else if (!squelched) {                                          // a frequency error
            a.consume = uint32_t(N) - value;
            s.ffe += f_index;
        } else {                                                          // just noise
            a.consume = uint32_t(N);
            s.ffe = 0.0f;
            s.fine_index = 0;
        }
    } break;
    case RX_DOWNCHIRP0: {                                                 // :242-251
        s.phase = RX_DOWNCHIRP1;
        a.consume = uint32_t(N);
        int error = int(value);
        if (value > unsigned(N / 2)) error -= N;
        s.freq_error = error;
    } break;
    case RX_DOWNCHIRP1: {                                                 // :256-271
        s.phase = RX_QUARTERCHIRP;
        a.consume = uint32_t(N);
        s.table = LORAB200_TABLE_UP;
        int error = int(value);
        if (value > unsigned(N / 2)) error -= N;
        s.freq_error = (s.freq_error + error) / 2;
        s.sig_error = s.freq_error;
        s.sig_power = power;
        s.sig_snr = snr;
    } break;
    case RX_QUARTERCHIRP: {                                               // :276-285
        s.phase = RX_DATASYMBOLS;
        a.consume = uint32_t(N / 4 + s.freq_error / 2);
        s.ffe += float(s.freq_error / 2);
        s.sym_count = 0;
    } break;
    default: {                                                            // RX_DATASYMBOLS  :290-312
        a.consume = uint32_t(N);
        a.emit_symbol = int32_t(int16_t(value));
        s.sym_count++;
        if (s.sym_count >= c.mtu || squelched) {
            a.post_packet = 1;
            s.ffe = 0.0f;
            s.phase = RX_FRAMESYNC;
        }
    } break;
    }
    s.prev_value = int32_t(int16_t(value));                               // :327  short _prevValue
    (void)power;
    return a;
}

} // namespace lorab200
