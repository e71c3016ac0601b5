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
// The state machine itself stays on the host and is a restatement of LoRaDemod.cpp:173-312,
// including its quirks (the sync-word look-ahead overwrites power/powerAvg/fIndex; the
// quarter-chirp skip uses integer halving).  Differences, all deliberate:
//   * _finefreqError, _prevValue, _freqError, _symCount are zero-initialised (the reference
//     reads _finefreqError uninitialised, LoRaDemod.cpp:160 / :386-392);
//   * work() may consume many windows per call; debug ports then carry the concatenation
//     of what the reference would have produced over the same number of calls;
//   * two extra setters: setDebugPorts(bool) (default true; false skips the raw/dec/fft
//     dumps and their D2H traffic) and setMaxBatch(n).
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
        _maxBatch(4096),
        _state(STATE_FRAMESYNC),
        _table(LORAB200_TABLE_UP),
        _symCount(0),
        _prevValue(0),
        _freqError(0),
        _fineTuneIndex(0),
        _finefreqError(0.0f)
    {
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, setSync));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, setThreshold));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, setMTU));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, setDebugPorts));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, setMaxBatch));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, getKernelLaunches));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, getDemodCalls));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDemodB200, getDemodWindows));
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
    unsigned long long getKernelLaunches(void) const { return lorab200_kernel_launches(_ctx); }
    unsigned long long getDemodCalls(void) const { return _demodCalls; } // lorab200_demod_host calls so far
    unsigned long long getDemodWindows(void) const { return _demodWindows; } // windows submitted (incl. unused speculation)

    void activate(void)
    {
        _state = STATE_FRAMESYNC;
        _table = LORAB200_TABLE_UP;
        _symCount = 0;
        _prevValue = 0;
        _freqError = 0;
        _fineTuneIndex = 0;
        _finefreqError = 0.0f;
    }

    void work(void)
    {
        auto inPort = this->input(0);
        const size_t avail = inPort->elements();
        if (avail < N*2) return;
        const auto inBuff = inPort->buffer().as<const std::complex<float> *>();

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

        while (avail - off >= N*2 and rawRoom - off >= N*2 and fftRoom - fftOut >= N)
        {
            const Window &win = this->demodAt(inBuff, avail, off, 0);
            const size_t value = size_t(win.res.index);
            float power = win.res.power;
            float powerAvg = win.res.power_avg;
            float fIndex = win.res.f_index;
            const float snr = power - powerAvg;
            const bool squelched = (snr < _thresh);

            //the dechirp loop advanced the fine tune index over the window (LoRaDemod.cpp:160-162)
            const int ftAtWindowStart = _fineTuneIndex;
            _fineTuneIndex = win.idxAfter;

            if (_debugPorts)
            {
                std::memcpy(decBuff + off, win.dec, N*sizeof(std::complex<float>));
                std::memcpy(fftBuff + fftOut, win.fft, N*sizeof(std::complex<float>));
            }

            size_t total = 0;
            std::string id;
            switch (_state)
            {
            ////////////////////////////////////////////////////////////////
            case STATE_FRAMESYNC:
            ////////////////////////////////////////////////////////////////
            {
                //format as observed from inspecting RN2483 (LoRaDemod.cpp:183-185)
                const bool syncd = not squelched and (_prevValue+4)/8 == 0;
                const bool match0 = (value+4)/8 == unsigned(_sync>>4);
                bool match1 = false;

                //the look-ahead window continues the fine tune walk from where the first
                //window left it, on a copy (LoRaDemod.cpp:191-203); its detect() call
                //overwrites power, powerAvg and fIndex
                if (syncd and match0)
                {
                    (void)ftAtWindowStart;
                    const Window &win1 = this->demodAt(inBuff, avail, off, 1);
                    power = win1.res.power;
                    powerAvg = win1.res.power_avg;
                    fIndex = win1.res.f_index;
                    match1 = (size_t(win1.res.index)+4)/8 == unsigned(_sync & 0xf);
                    if (_debugPorts) std::memcpy(decBuff + off + N, win1.dec, N*sizeof(std::complex<float>));
                }

                if (syncd and match0 and match1)
                {
                    total = 2*N;
                    _state = STATE_DOWNCHIRP0;
                    _table = LORAB200_TABLE_DOWN;
                    id = "SYNC";
                }
                //otherwise its a frequency error
                else if (not squelched)
                {
                    total = N - value;
                    _finefreqError += fIndex;
                    id = "P " + fixed4(fIndex);
                }
                //just noise
                else
                {
                    total = N;
                    _finefreqError = 0;
                    _fineTuneIndex = 0;
                }
            } break;

            ////////////////////////////////////////////////////////////////
            case STATE_DOWNCHIRP0:
            ////////////////////////////////////////////////////////////////
            {
                _state = STATE_DOWNCHIRP1;
                total = N;
                id = "DC";
                int error = int(value);
                if (value > N/2) error -= int(N);
                _freqError = error;
            } break;

            ////////////////////////////////////////////////////////////////
            case STATE_DOWNCHIRP1:
            ////////////////////////////////////////////////////////////////
            {
                _state = STATE_QUARTERCHIRP;
                total = N;
                _table = LORAB200_TABLE_UP;
                _outSymbols = Pothos::BufferChunk(typeid(int16_t), _mtu);

                int error = int(value);
                if (value > N/2) error -= int(N);
                _freqError = (_freqError + error)/2;

                this->emitSignal("error", _freqError);
                this->emitSignal("power", power);
                this->emitSignal("snr", snr);
            } break;

            ////////////////////////////////////////////////////////////////
            case STATE_QUARTERCHIRP:
            ////////////////////////////////////////////////////////////////
            {
                _state = STATE_DATASYMBOLS;
                total = N/4 + (_freqError / 2);
                _finefreqError += (_freqError / 2);
                _symCount = 0;
                id = "QC";
            } break;

            ////////////////////////////////////////////////////////////////
            case STATE_DATASYMBOLS:
            ////////////////////////////////////////////////////////////////
            {
                total = N;
                _outSymbols.as<int16_t *>()[_symCount++] = int16_t(value);
                if (_symCount >= _mtu or squelched)
                {
                    Pothos::Packet pkt;
                    pkt.payload = _outSymbols;
                    pkt.payload.length = _symCount*sizeof(int16_t);
                    this->output(0)->postMessage(pkt);
                    _finefreqError = 0;
                    _state = STATE_FRAMESYNC;
                }
                id = "S" + std::to_string(_symCount) + " " + fixed4(fIndex);
            } break;
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
            _prevValue = short(value);
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

    //! input slabs large enough for a batch of windows (the reference asks for 2N, LoRaDemod.cpp:348-358)
    Pothos::BufferManager::Sptr getInputBufferManager(const std::string &name, const std::string &domain)
    {
        if (name == "0")
        {
            Pothos::BufferManagerArgs args;
            args.bufferSize = std::max(args.bufferSize, N*(kDebugWindows+2)*sizeof(std::complex<float>));
            return Pothos::BufferManager::make("generic", args);
        }
        return Pothos::Block::getInputBufferManager(name, domain);
    }

private:
    enum { kDebugWindows = 64, kIdleSpecMin = 16 };

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
        int idxHere = _fineTuneIndex;          //index at the first sample of the requested window
        //the look-ahead window (ahead == 1) is only asked for after _fineTuneIndex has
        //been advanced over the first window: idxHere is already right in both cases
        const bool hit = c.count > 0 and at >= c.off and (at - c.off) % N == 0 and (at - c.off)/N < c.count
            and c.table == _table and c.ffe == _finefreqError
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
                if (_state == STATE_DATASYMBOLS)
                {
                    count = std::min<size_t>(maxWin, std::max<size_t>(1, _mtu - _symCount));
                }
                else if (_state == STATE_FRAMESYNC and _finefreqError == 0.0f and idxHere == 0)
                {
                    count = std::min(maxWin, _idleSpec);  //idle: every window starts from a reset fine tune state
                    resetEach = true;
                }
            }
            c.off = at; c.count = count; c.table = _table; c.idx0 = idxHere; c.ffe = _finefreqError;
            c.resetEach = resetEach;
            c.res.resize(count);
            c.idxAfter.resize(count);
            if (_debugPorts) { c.dec.resize(count*N); c.fft.resize(count*N); }
            lorab200_fine fine;
            fine.index = idxHere;
            fine.step = _finefreqError * _fineSteps;
            ++_demodCalls;
            _demodWindows += count;
            const int st = lorab200_demod_host(_ctx, reinterpret_cast<const float *>(in + at), count, N, _table, &fine,
                c.res.data(), _debugPorts ? reinterpret_cast<float *>(c.fft.data()) : nullptr,
                _debugPorts ? reinterpret_cast<float *>(c.dec.data()) : nullptr);
            if (st != LORAB200_OK)
            {
                throw Pothos::Exception("LoRaDemodB200::work()", std::string(lorab200_strerror(st)) + ": " + lorab200_last_error(_ctx));
            }
            int idx = idxHere;
            for (size_t k = 0; k < count; k++)
            {
                idx = resetEach ? 0 : this->rollIndex(idx, _finefreqError);
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
    size_t _idleSpec = kIdleSpecMin;        //current idle speculation depth (windows)
    unsigned char _sync;
    float _thresh;
    size_t _mtu;
    bool _debugPorts;
    size_t _maxBatch;
    Pothos::OutputPort *_rawPort;
    Pothos::OutputPort *_decPort;
    Pothos::OutputPort *_fftPort;

    //state
    enum LoraDemodState
    {
        STATE_FRAMESYNC,
        STATE_DOWNCHIRP0,
        STATE_DOWNCHIRP1,
        STATE_QUARTERCHIRP,
        STATE_DATASYMBOLS,
    };
    LoraDemodState _state;
    int _table;
    size_t _symCount;
    Pothos::BufferChunk _outSymbols;
    short _prevValue;
    int _freqError;
    int _fineTuneIndex;
    float _finefreqError;
};

static Pothos::BlockRegistry registerLoRaDemodB200(
    LORAB200_REGISTRY_PATH, &LoRaDemodB200::make);
