// LoRaDemodB200.cpp -- drop-in `/lora/lora_demod(sf)` block whose hot loop runs on a B200.
//
// Same registry path, ports, setters, signals, labels and frame-sync state machine as the
// reference block (LoRaDemod.cpp:65-396); the per-window work -- dechirp
// (LoRaDemod.cpp:157-166, twin :192-202) + LoRaDetector::detect (LoRaDetector.hpp:29-64)
// -- is done by the fused CUDA kernel through the C ABI (include/lorab200.h,
// lorab200_demod_host), for as many windows per work() call as the state machine allows:
//
//   * STATE_DATASYMBOLS: window positions are known in advance (stride N, constant
//     _finefreqError, LoRaDemod.cpp:289 and the commented-out update at :308), so all
//     symbols of the packet that are already in the input buffer go to the GPU in one call;
//   * STATE_FRAMESYNC while idle (squelched, error terms reset): windows are consumed N at
//     a time with a zero fine-tune state (LoRaDemod.cpp:229-235), so the whole buffer is
//     demodulated speculatively in one call and walked until the first un-squelched window;
//   * everything else (preamble re-timing `total = N - value`, sync word, down-chirps,
//     quarter chirp) depends on the previous result and is done one window at a time.
//
// The state machine stays on the host: it is the transition function of csrc/rx_fsm.h
// (rx_needs_second / rx_step, the restatement of LoRaDemod.cpp:173-327 that the on-device
// receiver lorab200_rx_dev runs per channel too), so the block and the device receiver cannot
// drift apart.  This file adds what a Pothos block owns around it: ports, packets, labels,
// signals, and the batching of GPU calls.  Differences from the reference, all deliberate:
//   * _finefreqError, _prevValue, _freqError, _symCount are zero-initialised (the reference
//     reads _finefreqError uninitialised, LoRaDemod.cpp:160 / :386-392);
//   * work() may consume many windows per call; debug ports then carry the concatenation
//     of what the reference would have produced over the same number of calls;
//   * two extra setters: setDebugPorts(bool) (default true; false skips the raw/dec/fft
//     dumps and their D2H traffic) and setMaxBatch(n);
//   * with the debug ports off, work() does not walk windows on the host at all: it hands the
//     whole input buffer to the on-device receiver (lorab200_rx_host -> rx_batch.cuh: the same
//     state machine, one launch per work() call, speculative batches of windows), posts the
//     packets it returns and consumes what it consumed.  The state it carries between calls is
//     this block's _st.  setDeviceReceiver(false) keeps the host walk (this file's original path).
//
// Builds against <Pothos/Framework.hpp> -- the real one, or ministream's stand-in.
#include <Pothos/Framework.hpp>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/lorab200.h"
#include "../../csrc/rx_fsm.h"

#ifndef LORAB200_REGISTRY_PATH
#define LORAB200_REGISTRY_PATH "/lora/lora_demod"
#endif

/***********************************************************************
 * |PothosDoc LoRa Demod (B200)
 *
 * Demodulate LoRa packets from a complex sample stream into symbols.
 * Drop-in replacement for the LoRa Demod block: the dechirp, FFT and peak
 * detection of every symbol window run in one fused CUDA kernel on a B200.
 *
 * <h2>Input format</h2>
 * The input port 0 accepts a complex sample stream of modulated chirps
 * received at the specified bandwidth and carrier frequency.
 *
 * <h2>Output format</h2>
 * The output port 0 produces a packet containing demodulated symbols.
 * The format of the packet payload is a buffer of unsigned shorts.
 *
 * <h2>Debug ports raw, dec, fft</h2>
 * The input signal, the dechirped signal and the FFT of every processed window,
 * annotated with labels at the synchronization points.
 *
 * |category /LoRa
 * |keywords lora
 *
 * |param sf[Spread factor] Each symbol occupies 2^SF samples.
 * |default 10
 *
 * |param sync[Sync word] The 2-nibble, 2-symbol sync value.
 * |default 0x12
 *
 * |param thresh[Threshold] The minimum required level in dB for the detector.
 * |units dB
 * |default -30.0
 *
 * |param mtu[Symbol MTU] Produce MTU at most symbols after sync is found.
 * |units symbols
 * |default 256
 *
 * |factory /lora/lora_demod(sf)
 * |setter setSync(sync)
 * |setter setThreshold(thresh)
 * |setter setMTU(mtu)
 **********************************************************************/
//! Slabs from the library's page-locked allocator (cudaMallocHost behind the C ABI).  With ministream's
//! Pothos surface a manager is asked for slabs through allocate() / release(); INTEGRATION.md shows the
//! same manager against the real Pothos::BufferManager interface (init / SharedBuffer slabs).
class PinnedBufferManager : public Pothos::BufferManager
{
public:
    void *allocate(size_t bytes) override { return lorab200_alloc_pinned(bytes); }
    void release(void *p) override { lorab200_free_pinned(p); }
    bool custom() const override { return true; }
};

class LoRaDemodB200 : public Pothos::Block
{
public:
    LoRaDemodB200(const size_t sf):
        N(size_t(1) << sf),
        _sf(int(sf)),
        _fineSteps(128),
        _ctx(nullptr),
        _sync(0x12),
        _thresh(-30.0f),
        _mtu(256),
        _debugPorts(true),
        _deviceReceiver(true),
        _maxBatch(4096)
    {
        lorab200::rx_reset(_st);
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, setSync));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, setThreshold));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, setMTU));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, setDebugPorts));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, setMaxBatch));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, setDeviceReceiver));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, getKernelLaunches));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, getDemodCalls));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, getDemodWindows));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, getPinnedInputs));
        this->setupInput(0, typeid(std::complex<float>));
        this->setupOutput(0);
        this->setupOutput("raw", typeid(std::complex<float>));
        this->setupOutput("dec", typeid(std::complex<float>));
        this->setupOutput("fft", typeid(std::complex<float>));

        this->registerSignal("error");
        this->registerSignal("power");
        this->registerSignal("snr");

        //a window is only processed while two symbols are available (LoRaDemod.cpp:90,148)
        this->input(0)->setReserve(N*2);

        _rawPort = this->output("raw");
        _decPort = this->output("dec");
        _fftPort = this->output("fft");

        int device = 0;
        if (const char *env = std::getenv("LORAB200_DEVICE")) device = std::atoi(env);
        const int st = lorab200_create(&_ctx, device, int(sf), LORAB200_FLAG_DEFAULT);
        if (st != LORAB200_OK)
        {
            throw Pothos::Exception("LoRaDemodB200(" + std::to_string(sf) + ")",
                std::string(lorab200_strerror(st)) + ": " + lorab200_last_error(nullptr));
        }
    }

    ~LoRaDemodB200(void)
    {
        lorab200_destroy(_ctx);
    }

    static Block *make(const size_t sf)
    {
        return new LoRaDemodB200(sf);
    }

    void setSync(const unsigned char sync) { _sync = sync; }
    void setThreshold(const double thresh_dB) { _thresh = thresh_dB; }
    void setMTU(const size_t mtu) { _mtu = mtu; }
    void setDebugPorts(const bool enable) { _debugPorts = enable; }
    void setMaxBatch(const size_t n) { _maxBatch = std::max<size_t>(1, n); }
    void setDeviceReceiver(const bool enable) { _deviceReceiver = enable; }
    unsigned long long getKernelLaunches(void) const { return lorab200_kernel_launches(_ctx); }
    unsigned long long getDemodCalls(void) const { return _demodCalls; } // lorab200_demod_host calls so far
    unsigned long long getDemodWindows(void) const { return _demodWindows; } // windows submitted (incl. unused speculation)
    unsigned long long getPinnedInputs(void) const { return _pinnedInputs; } // work() calls whose input buffer was page-locked

    void activate(void)
    {
        lorab200::rx_reset(_st);
        _sigEmitted = false;
    }

    //! debug ports off: the whole input buffer goes through the on-device receiver in one call
    void workReceiver(void)
    {
        auto inPort = this->input(0);
        const size_t avail = inPort->elements();
        if (avail < N*2) return;
        const float *iq = reinterpret_cast<const float *>(inPort->buffer().as<const std::complex<float> *>());
        _pinnedInputs += lorab200_is_pinned(iq);
        const size_t n = std::min(avail, _maxBatch*N);
        if (_rxMtu != _mtu)
        {
            //setMTU between packets: re-shape the scratch (a packet under construction keeps its symbols)
            _rxPartial.resize(_mtu, 0);
            _rxSymbols.assign(size_t(kRxMaxPackets)*_mtu, 0);
            _rxMtu = _mtu;
        }
        static_assert(sizeof(lorab200_rx_state) == sizeof(lorab200::RxState), "the C struct mirrors RxState");
        lorab200_rx_state st;
        std::memcpy(&st, &_st, sizeof(st));
        lorab200_rx_packet info[kRxMaxPackets];
        int32_t npk = 0;
        uint64_t consumed = 0;
        ++_demodCalls;
        _demodWindows += n/N;
        const int rc = lorab200_rx_host(_ctx, &iq, 1, n, _sync, _thresh, _mtu, &st, _rxPartial.data(),
            _rxSymbols.data(), info, &npk, &consumed, kRxMaxPackets);
        if (rc != LORAB200_OK)
        {
            throw Pothos::Exception("LoRaDemodB200::work()", std::string(lorab200_strerror(rc)) + ": " + lorab200_last_error(_ctx));
        }
        std::memcpy(&_st, &st, sizeof(st));
        for (int32_t p = 0; p < npk; p++)
        {
            //the signals belong to the packet's DOWNCHIRP1 window (LoRaDemod.cpp:267-269): emitted when that
            //window is passed -- at the end of an earlier call for a packet that was still being received
            if (not (p == 0 and _sigEmitted))
            {
                this->emitSignal("error", int(info[p].error));
                this->emitSignal("power", info[p].power);
                this->emitSignal("snr", info[p].snr);
            }
            _sigEmitted = false;
            Pothos::Packet pkt;
            pkt.payload = Pothos::BufferChunk(typeid(int16_t), _mtu);
            std::memcpy(pkt.payload.as<int16_t *>(), _rxSymbols.data() + size_t(p)*_mtu, size_t(info[p].length)*sizeof(int16_t));
            pkt.payload.length = size_t(info[p].length)*sizeof(int16_t);
            this->output(0)->postMessage(pkt);
        }
        if ((_st.phase == lorab200::RX_QUARTERCHIRP or _st.phase == lorab200::RX_DATASYMBOLS) and not _sigEmitted)
        {
            this->emitSignal("error", int(_st.sig_error));
            this->emitSignal("power", _st.sig_power);
            this->emitSignal("snr", _st.sig_snr);
            _sigEmitted = true;
        }
        inPort->consume(size_t(consumed));
    }

    void work(void)
    {
        if (not _debugPorts and _deviceReceiver) return this->workReceiver();
        auto inPort = this->input(0);
        const size_t avail = inPort->elements();
        if (avail < N*2) return;
        const auto inBuff = inPort->buffer().as<const std::complex<float> *>();
        _pinnedInputs += lorab200_is_pinned(inBuff);

        //room on the debug ports bounds how many windows this call may process
        size_t rawRoom = size_t(-1), fftRoom = size_t(-1);
        std::complex<float> *rawBuff = nullptr, *decBuff = nullptr, *fftBuff = nullptr;
        if (_debugPorts)
        {
            rawBuff = _rawPort->buffer().as<std::complex<float> *>();
            decBuff = _decPort->buffer().as<std::complex<float> *>();
            fftBuff = _fftPort->buffer().as<std::complex<float> *>();
            rawRoom = std::min(_rawPort->elements(), _decPort->elements());
            fftRoom = _fftPort->elements();
        }

        size_t off = 0;       //input samples consumed so far in this call
        size_t fftOut = 0;    //elements produced on the fft port in this call
        _cache.count = 0;

        lorab200::RxCfg cfg;
        cfg.n = int(N); cfg.sync = int(_sync); cfg.thresh = _thresh; cfg.mtu = int(_mtu);

        while (avail - off >= N*2 and rawRoom - off >= N*2 and fftRoom - fftOut >= N)
        {
            //the window at the current offset, and -- when it looks like the first sync word -- the next one
            const Window &win = this->demodAt(inBuff, avail, off, 0);
            const lorab200_result r0 = win.res;
            const int idxAfter0 = win.idxAfter;
            if (_debugPorts)
            {
                std::memcpy(decBuff + off, win.dec, N*sizeof(std::complex<float>));
                std::memcpy(fftBuff + fftOut, win.fft, N*sizeof(std::complex<float>));
            }
            const bool second = lorab200::rx_needs_second(cfg, _st, r0);
            lorab200_result r1 = r0;
            if (second)
            {
                //demodAt continues the fine tune walk from where the first window left it
                _st.fine_index = idxAfter0;
                const Window &win1 = this->demodAt(inBuff, avail, off, 1);
                r1 = win1.res;
                if (_debugPorts) std::memcpy(decBuff + off + N, win1.dec, N*sizeof(std::complex<float>));
            }

            const int before = _st.phase;
            const bool squelched = (r0.power - r0.power_avg) < _thresh;
            const lorab200::RxAction act = lorab200::rx_step(cfg, _st, r0, idxAfter0, second, r1);
            const size_t total = act.consume;

            //what the block owns around a transition: the packet, the signals, the labels
            std::string id;
            switch (before)
            {
            case lorab200::RX_FRAMESYNC:
                if (_st.phase == lorab200::RX_DOWNCHIRP0) id = "SYNC";
                else if (not squelched) id = "P " + fixed4(second ? r1.f_index : r0.f_index);
                break;
            case lorab200::RX_DOWNCHIRP0:
                id = "DC";
                break;
            case lorab200::RX_DOWNCHIRP1:
                _outSymbols = Pothos::BufferChunk(typeid(int16_t), _mtu);
                this->emitSignal("error", int(_st.sig_error));
                this->emitSignal("power", _st.sig_power);
                this->emitSignal("snr", _st.sig_snr);
                break;
            case lorab200::RX_QUARTERCHIRP:
                id = "QC";
                break;
            default: //RX_DATASYMBOLS
                _outSymbols.as<int16_t *>()[_st.sym_count - 1] = int16_t(act.emit_symbol);
                if (act.post_packet)
                {
                    Pothos::Packet pkt;
                    pkt.payload = _outSymbols;
                    pkt.payload.length = size_t(_st.sym_count)*sizeof(int16_t);
                    this->output(0)->postMessage(pkt);
                }
                id = "S" + std::to_string(_st.sym_count) + " " + fixed4(r0.f_index);
                break;
            }

            if (_debugPorts)
            {
                std::memcpy(rawBuff + off, inBuff + off, total*sizeof(std::complex<float>));
                if (not id.empty())
                {
                    _rawPort->postLabel(Pothos::Label(id, Pothos::Object(), off));
                    _decPort->postLabel(Pothos::Label(id, Pothos::Object(), off));
                    _fftPort->postLabel(Pothos::Label(id, Pothos::Object(), fftOut));
                }
                fftOut += N;
            }
            off += total;
        }

        inPort->consume(off);
        if (_debugPorts)
        {
            _rawPort->produce(off);
            _decPort->produce(off);
            _fftPort->produce(fftOut);
        }
    }

    //! debug ports carry up to a batch of windows per work() call
    Pothos::BufferManager::Sptr getOutputBufferManager(const std::string &name, const std::string &domain)
    {
        if (name == "raw" or name == "dec" or name == "fft")
        {
            this->output(name)->setReserve(N * 2);
            Pothos::BufferManagerArgs args;
            args.bufferSize = N*(kDebugWindows+2)*sizeof(std::complex<float>);
            return Pothos::BufferManager::make("generic", args);
        }
        return Pothos::Block::getOutputBufferManager(name, domain);
    }

    //! Input slabs large enough for a batch of windows (the reference asks for 2N, LoRaDemod.cpp:348-358)
    //! and PAGE-LOCKED: the upstream block then writes its samples straight into memory the GPU's copy
    //! engine reads at PCIe speed, and lorab200_demod_host / lorab200_rx_host skip their staging ring
    //! (round 1: 11 Msym/s through pageable port buffers, 48 through page-locked ones).
    Pothos::BufferManager::Sptr getInputBufferManager(const std::string &name, const std::string &domain)
    {
        if (name == "0")
        {
            auto mgr = std::make_shared<PinnedBufferManager>();
            mgr->name = "lorab200_pinned";
            mgr->args.bufferSize = std::max(mgr->args.bufferSize, N*(kDebugWindows+2)*sizeof(std::complex<float>));
            return mgr;
        }
        return Pothos::Block::getInputBufferManager(name, domain);
    }

private:
    enum { kDebugWindows = 64, kIdleSpecMin = 16, kRxMaxPackets = 16 };

    struct Window
    {
        lorab200_result res;
        int idxAfter;                       //_fineTuneIndex after the window's N samples
        const std::complex<float> *dec;     //dechirped samples (debug ports on)
        const std::complex<float> *fft;     //FFT bins (debug ports on)
    };

    //results of one GPU call: `count` windows at stride N from sample `off`, all
    //demodulated under the same assumed state (table, fine index at the first sample,
    //fine step, per-window index reset)
    struct Cache
    {
        size_t off = 0, count = 0;
        int table = 0, idx0 = 0;
        float ffe = 0;
        bool resetEach = false;
        std::vector<lorab200_result> res;
        std::vector<int> idxAfter;
        std::vector<std::complex<float>> dec, fft;
        Window win;
    } _cache;

    //LoRaDemod.cpp:160-162 advanced over a whole window, in closed form (lorab200_fine_advance)
    int rollIndex(int idx, const float ffe) const
    {
        const float step = ffe * _fineSteps;
        if (step == 0.0f) return idx;
        const int32_t r = lorab200_fine_advance(_sf, idx, step, N);
        if (r < 0) throw Pothos::Exception("LoRaDemodB200::work()", "fine tune state out of range");
        return r;
    }

    //! result for the window that starts `ahead` windows after sample `off` under the current state
    const Window &demodAt(const std::complex<float> *in, const size_t avail, const size_t off, const size_t ahead)
    {
        Cache &c = _cache;
        const size_t at = off + ahead*N;
        int idxHere = _st.fine_index;          //index at the first sample of the requested window
        //the look-ahead window (ahead == 1) is only asked for after _fineTuneIndex has
        //been advanced over the first window: idxHere is already right in both cases
        const bool hit = c.count > 0 and at >= c.off and (at - c.off) % N == 0 and (at - c.off)/N < c.count
            and c.table == _st.table and c.ffe == _st.ffe
            and (c.resetEach ? idxHere == 0 : this->cacheIdxAt((at - c.off)/N) == idxHere);
        if (not hit)
        {
            //adapt the idle speculation depth to how much of the last idle batch was used: double
            //after a batch that was consumed to the end, shrink to what was used otherwise, so
            //that a long silence is demodulated in a few large calls while a busy channel does
            //not pay for thousands of windows that the next preamble invalidates
            if (c.count > 0 and c.resetEach and at >= c.off)
            {
                const size_t used = (at - c.off + N - 1)/N;
                if (used >= c.count) _idleSpec = std::min(_idleSpec*2, _maxBatch);
                else _idleSpec = std::max<size_t>(kIdleSpecMin, std::min(_idleSpec, 2*used));
            }
            //how many windows may be demodulated speculatively from here?
            size_t count = 1;
            const size_t fit = (avail - at)/N;              //whole windows in the buffer from `at`
            bool resetEach = false;
            if (ahead == 0 and fit >= 2)
            {
                const size_t maxWin = std::min<size_t>(fit - 1, _debugPorts ? size_t(kDebugWindows) : _maxBatch);
                if (_st.phase == lorab200::RX_DATASYMBOLS)
                {
                    count = std::min<size_t>(maxWin, std::max<size_t>(1, _mtu - size_t(_st.sym_count)));
                }
                else if (_st.phase == lorab200::RX_FRAMESYNC and _st.ffe == 0.0f and idxHere == 0)
                {
                    count = std::min(maxWin, _idleSpec);  //idle: every window starts from a reset fine tune state
                    resetEach = true;
                }
            }
            c.off = at; c.count = count; c.table = _st.table; c.idx0 = idxHere; c.ffe = _st.ffe;
            c.resetEach = resetEach;
            c.res.resize(count);
            c.idxAfter.resize(count);
            if (_debugPorts) { c.dec.resize(count*N); c.fft.resize(count*N); }
            lorab200_fine fine;
            fine.index = idxHere;
            fine.step = _st.ffe * _fineSteps;
            ++_demodCalls;
            _demodWindows += count;
            const int st = lorab200_demod_host(_ctx, reinterpret_cast<const float *>(in + at), count, N, _st.table, &fine,
                c.res.data(), _debugPorts ? reinterpret_cast<float *>(c.fft.data()) : nullptr,
                _debugPorts ? reinterpret_cast<float *>(c.dec.data()) : nullptr);
            if (st != LORAB200_OK)
            {
                throw Pothos::Exception("LoRaDemodB200::work()", std::string(lorab200_strerror(st)) + ": " + lorab200_last_error(_ctx));
            }
            int idx = idxHere;
            for (size_t k = 0; k < count; k++)
            {
                idx = resetEach ? 0 : this->rollIndex(idx, _st.ffe);
                c.idxAfter[k] = idx;
            }
        }
        const size_t k = (at - c.off)/N;
        c.win.res = c.res[k];
        c.win.idxAfter = c.idxAfter[k];
        c.win.dec = _debugPorts ? c.dec.data() + k*N : nullptr;
        c.win.fft = _debugPorts ? c.fft.data() + k*N : nullptr;
        return c.win;
    }

    int cacheIdxAt(const size_t k) const
    {
        return k == 0 ? _cache.idx0 : _cache.idxAfter[k-1];
    }

    static std::string fixed4(const float v)
    {
        std::stringstream stream;
        stream.precision(4);
        stream << std::fixed << v;
        return stream.str();
    }

    //configuration
    const size_t N;
    const int _sf;
    const size_t _fineSteps;
    lorab200_ctx *_ctx;
    unsigned long long _demodCalls = 0;
    unsigned long long _demodWindows = 0;
    unsigned long long _pinnedInputs = 0;
    size_t _idleSpec = kIdleSpecMin;        //current idle speculation depth (windows)
    unsigned char _sync;
    float _thresh;
    size_t _mtu;
    bool _debugPorts;
    bool _deviceReceiver;
    size_t _maxBatch;
    //device-receiver path: scratch between calls (the packet under construction, returned packets)
    std::vector<int16_t> _rxPartial, _rxSymbols;
    size_t _rxMtu = 0;
    bool _sigEmitted = false;               //the signals of the packet under construction are out already
    Pothos::OutputPort *_rawPort;
    Pothos::OutputPort *_decPort;
    Pothos::OutputPort *_fftPort;

    //state: the reference block's private members (LoRaDemod.cpp:362-392) as one struct, plus the packet
    lorab200::RxState _st;
    Pothos::BufferChunk _outSymbols;
};

static Pothos::BlockRegistry registerLoRaDemodB200(
    LORAB200_REGISTRY_PATH, &LoRaDemodB200::make);

/***********************************************************************
 * |PothosDoc LoRa Demod, many channels (B200)
 *
 * K independent LoRa channels of the same spread factor behind ONE block: input ports "0" .. "K-1"
 * (complex sample streams), output ports "0" .. "K-1" (symbol packets, as the LoRa Demod block posts
 * them).  One work() call hands every port's buffer to the on-device receiver in ONE launch
 * (lorab200_rx_host -> one CTA per channel); each channel keeps its own frame-sync state, exactly as K
 * separate LoRa Demod blocks would (LoRaDemod.cpp:145-327 per channel), without K threads and K launches.
 * No debug ports.  Signals error / power / snr carry the channel number as a first argument.
 *
 * |category /LoRa
 * |keywords lora
 *
 * |param sf[Spread factor] |default 10
 * |param channels[Channels] Number of input / output port pairs. |default 8
 * |param sync[Sync word] |default 0x12
 * |param thresh[Threshold] |units dB |default -30.0
 * |param mtu[Symbol MTU] |units symbols |default 256
 *
 * |factory /lora/lora_demod_multi(sf, channels)
 * |setter setSync(sync)
 * |setter setThreshold(thresh)
 * |setter setMTU(mtu)
 **********************************************************************/
class LoRaDemodMultiB200 : public Pothos::Block
{
public:
    LoRaDemodMultiB200(const size_t sf, const size_t channels):
        N(size_t(1) << sf), K(channels), _ctx(nullptr), _sync(0x12), _thresh(-30.0f), _mtu(256), _maxBatch(4096)
    {
        if (K == 0 or K > 4096) throw Pothos::InvalidArgumentException("LoRaDemodMultiB200", "channels must be 1..4096");
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodMultiB200, setSync));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodMultiB200, setThreshold));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodMultiB200, setMTU));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodMultiB200, setMaxBatch));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodMultiB200, getKernelLaunches));
        for (size_t c = 0; c < K; c++)
        {
            this->setupInput(c, typeid(std::complex<float>));
            this->setupOutput(c);
            this->input(c)->setReserve(N*2);
        }
        this->registerSignal("error");
        this->registerSignal("power");
        this->registerSignal("snr");
        int device = 0;
        if (const char *env = std::getenv("LORAB200_DEVICE")) device = std::atoi(env);
        const int st = lorab200_create(&_ctx, device, int(sf), LORAB200_FLAG_DEFAULT);
        if (st != LORAB200_OK)
            throw Pothos::Exception("LoRaDemodMultiB200(" + std::to_string(sf) + ")",
                std::string(lorab200_strerror(st)) + ": " + lorab200_last_error(nullptr));
        _state.resize(K);
        _sigEmitted.assign(K, 0);
    }

    ~LoRaDemodMultiB200(void) { lorab200_destroy(_ctx); }

    static Block *make(const size_t sf, const size_t channels) { return new LoRaDemodMultiB200(sf, channels); }

    void setSync(const unsigned char sync) { _sync = sync; }
    void setThreshold(const double thresh_dB) { _thresh = thresh_dB; }
    void setMTU(const size_t mtu) { _mtu = mtu; }
    void setMaxBatch(const size_t n) { _maxBatch = std::max<size_t>(1, n); }
    unsigned long long getKernelLaunches(void) const { return lorab200_kernel_launches(_ctx); }

    void activate(void)
    {
        std::memset(_state.data(), 0, _state.size()*sizeof(lorab200_rx_state));   //all zero == rx_reset
        _sigEmitted.assign(K, 0);
    }

    void work(void)
    {
        //one launch over the channels that have at least two symbols buffered (LoRaDemod.cpp:148), for as
        //many samples as the shortest of them holds; the scheduler calls again while any port has more
        if (_rxMtu != _mtu)
        {
            _partial.assign(K*_mtu, 0);
            _symbols.assign(K*size_t(kRxMaxPackets)*_mtu, 0);
            _rxMtu = _mtu;
        }
        std::vector<size_t> run;
        size_t n = _maxBatch*N;
        for (size_t c = 0; c < K; c++)
        {
            const size_t have = this->input(c)->elements();
            if (have < N*2) continue;
            run.push_back(c);
            n = std::min(n, have);
        }
        const size_t R = run.size();
        if (R == 0) return;
        std::vector<const float *> iq(R);
        std::vector<lorab200_rx_state> state(R);
        std::vector<int16_t> partial(R*_mtu);
        for (size_t i = 0; i < R; i++)
        {
            const size_t c = run[i];
            iq[i] = reinterpret_cast<const float *>(this->input(c)->buffer().as<const std::complex<float> *>());
            state[i] = _state[c];
            std::memcpy(partial.data() + i*_mtu, _partial.data() + c*_mtu, _mtu*sizeof(int16_t));
        }
        std::vector<lorab200_rx_packet> info(R*size_t(kRxMaxPackets));
        std::vector<int32_t> npk(R);
        std::vector<uint64_t> consumed(R);
        const int rc = lorab200_rx_host(_ctx, iq.data(), R, n, _sync, _thresh, _mtu, state.data(), partial.data(),
            _symbols.data(), info.data(), npk.data(), consumed.data(), kRxMaxPackets);
        if (rc != LORAB200_OK)
            throw Pothos::Exception("LoRaDemodMultiB200::work()", std::string(lorab200_strerror(rc)) + ": " + lorab200_last_error(_ctx));
        for (size_t i = 0; i < R; i++)
        {
            const size_t c = run[i];
            _state[c] = state[i];
            std::memcpy(_partial.data() + c*_mtu, partial.data() + i*_mtu, _mtu*sizeof(int16_t));
            const lorab200_rx_packet *ci = info.data() + i*size_t(kRxMaxPackets);
            for (int32_t p = 0; p < npk[i]; p++)
            {
                if (not (p == 0 and _sigEmitted[c]))
                {
                    this->emitSignal("error", int(c), int(ci[p].error));
                    this->emitSignal("power", int(c), ci[p].power);
                    this->emitSignal("snr", int(c), ci[p].snr);
                }
                _sigEmitted[c] = 0;
                Pothos::Packet pkt;
                pkt.payload = Pothos::BufferChunk(typeid(int16_t), _mtu);
                std::memcpy(pkt.payload.as<int16_t *>(), _symbols.data() + (i*size_t(kRxMaxPackets) + size_t(p))*_mtu,
                    size_t(ci[p].length)*sizeof(int16_t));
                pkt.payload.length = size_t(ci[p].length)*sizeof(int16_t);
                this->output(c)->postMessage(pkt);
            }
            const lorab200_rx_state &st = _state[c];
            if ((st.phase == lorab200::RX_QUARTERCHIRP or st.phase == lorab200::RX_DATASYMBOLS) and not _sigEmitted[c])
            {
                this->emitSignal("error", int(c), int(st.sig_error));
                this->emitSignal("power", int(c), st.sig_power);
                this->emitSignal("snr", int(c), st.sig_snr);
                _sigEmitted[c] = 1;
            }
            this->input(c)->consume(size_t(consumed[i]));
        }
    }

    //! page-locked input slabs, large enough for a batch (as the single-channel block)
    Pothos::BufferManager::Sptr getInputBufferManager(const std::string &, const std::string &)
    {
        auto mgr = std::make_shared<PinnedBufferManager>();
        mgr->name = "lorab200_pinned";
        mgr->args.bufferSize = std::max(mgr->args.bufferSize, N*66*sizeof(std::complex<float>));
        return mgr;
    }

private:
    enum { kRxMaxPackets = 16 };
    const size_t N, K;
    lorab200_ctx *_ctx;
    unsigned char _sync;
    float _thresh;
    size_t _mtu, _maxBatch;
    size_t _rxMtu = 0;
    std::vector<lorab200_rx_state> _state;
    std::vector<int16_t> _partial, _symbols;
    std::vector<char> _sigEmitted;
};

#ifndef LORAB200_MULTI_REGISTRY_PATH
#define LORAB200_MULTI_REGISTRY_PATH "/lora/lora_demod_multi"
#endif
static Pothos::BlockRegistry registerLoRaDemodMultiB200(
    LORAB200_MULTI_REGISTRY_PATH, &LoRaDemodMultiB200::make);

