// LoRaCodecB200.cpp -- drop-in `/lora/lora_encoder()`, `/lora/lora_decoder()` and `/lora/lora_mod(sf)`
// blocks whose packet work runs on a B200 through the C ABI (include/lorab200.h):
//
//   LoRaEncoderB200  payload bytes -> symbols      lorab200_encode_host   (LoRaEncoder.cpp:84-91,137-228,244-245)
//   LoRaDecoderB200  symbols -> payload bytes      lorab200_decode_host   (LoRaDecoder.cpp:111-125,196-405,420-421)
//   LoRaModB200      symbols -> complex samples    lorab200_mod_host      (LoRaMod.cpp:65-70,120-229,280-281)
//
// Same registry paths (overridable for a side-by-side build), ports, setters, the decoder's
// `dropped` signal / getDropped probe, and the same behaviour per packet; what changes is that
// work() takes EVERY packet that is waiting on the input port and runs the batch in one launch --
// the per-packet integer work only pays on a GPU when packets arrive in numbers (a multi-channel
// gateway), which is exactly when a Pothos actor finds several messages queued.
//
// The device codec (csrc/codec.h) is bit-identical to the reference encoder / decoder for every
// setter combination (tests/test_codec.py, tests/test_gpu_decode.py); the modulator runs in
// LORAB200_MOD_REFERENCE mode (sample phases bit-identical to LoRaMod.cpp, samples within 2 float
// ulp) unless setIdeal(true) selects the closed-form chirps.
//
// Builds against <Pothos/Framework.hpp> -- the real one, or ministream's stand-in.
#include <Pothos/Framework.hpp>

#include <algorithm>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/lorab200.h"

#ifndef LORAB200_ENCODER_PATH
#define LORAB200_ENCODER_PATH "/lora/lora_encoder"
#endif
#ifndef LORAB200_DECODER_PATH
#define LORAB200_DECODER_PATH "/lora/lora_decoder"
#endif
#ifndef LORAB200_MOD_PATH
#define LORAB200_MOD_PATH "/lora/lora_mod"
#endif

namespace {

int deviceFromEnv(void)
{
    const char *env = std::getenv("LORAB200_DEVICE");
    return env ? std::atoi(env) : 0;
}

//! one library context per spreading factor the block has been asked for (the codec kernels read SF from it)
class CtxHolder
{
public:
    ~CtxHolder(void) { if (_ctx) lorab200_destroy(_ctx); }
    lorab200_ctx *get(const std::string &who, const size_t sf)
    {
        if (_ctx and _sf == sf) return _ctx;
        if (_ctx) { lorab200_destroy(_ctx); _ctx = nullptr; }
        const int st = lorab200_create(&_ctx, deviceFromEnv(), int(sf), LORAB200_FLAG_DEFAULT);
        if (st != LORAB200_OK)
            throw Pothos::Exception(who, std::string(lorab200_strerror(st)) + ": " + lorab200_last_error(nullptr));
        _sf = sf;
        return _ctx;
    }
private:
    lorab200_ctx *_ctx = nullptr;
    size_t _sf = 0;
};

int codingRate(const std::string &who, const std::string &cr)
{
    if (cr == "4/4") return 0;
    if (cr == "4/5") return 1;
    if (cr == "4/6") return 2;
    if (cr == "4/7") return 3;
    if (cr == "4/8") return 4;
    throw Pothos::InvalidArgumentException(who + "::setCodingRate(" + cr + ")", "unknown coding rate");
}

void check(const std::string &who, lorab200_ctx *ctx, const int st)
{
    if (st != LORAB200_OK)
        throw Pothos::Exception(who, std::string(lorab200_strerror(st)) + ": " + lorab200_last_error(ctx));
}

} // namespace

/***********************************************************************
 * |PothosDoc LoRa Encoder (B200)
 *
 * Encode bytes into LoRa modulation symbols.  Drop-in replacement for the LoRa Encoder block:
 * all packets waiting on the input port are encoded by one kernel launch on a B200.
 *
 * |category /LoRa
 * |keywords lora
 *
 * |param sf[Spread factor] The spread factor sets the bit width of the symbols.
 * |default 10
 * |param ppm[Symbol size] The size of the symbols in bits (0 = spread factor).
 * |default 0
 * |param cr[Coding Rate] "4/4" .. "4/8".
 * |default "4/8"
 * |param explicit Enable/disable explicit header mode.
 * |default true
 * |param crc Enable/disable the data CRC.
 * |default true
 * |param whitening Enable/disable whitening.
 * |default true
 *
 * |factory /lora/lora_encoder()
 * |setter setSpreadFactor(sf)
 * |setter setSymbolSize(ppm)
 * |setter setCodingRate(cr)
 * |setter enableExplicit(explicit)
 * |setter enableCrc(crc)
 * |setter enableWhitening(whitening)
 **********************************************************************/
class LoRaEncoderB200 : public Pothos::Block
{
public:
    LoRaEncoderB200(void): _sf(10), _ppm(0), _rdd(4), _explicit(true), _crc(true), _whitening(true)
    {
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaEncoderB200, setSpreadFactor));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaEncoderB200, setSymbolSize));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaEncoderB200, setCodingRate));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaEncoderB200, enableWhitening));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaEncoderB200, enableExplicit));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaEncoderB200, enableCrc));
        this->setupInput("0");
        this->setupOutput("0");
    }

    static Block *make(void) { return new LoRaEncoderB200(); }

    void setSpreadFactor(const size_t sf) { _sf = sf; }
    void setSymbolSize(const size_t ppm) { _ppm = ppm; }
    void setCodingRate(const std::string &cr) { _rdd = size_t(codingRate("LoRaEncoder", cr)); }
    void enableWhitening(const bool whitening) { _whitening = whitening; }
    void enableExplicit(const bool e) { _explicit = e; }
    void enableCrc(const bool crc) { _crc = crc; }

    void work(void)
    {
        auto inPort = this->input(0);
        auto outPort = this->output(0);
        if (not inPort->hasMessage()) return;
        const size_t PPM = (_ppm == 0) ? _sf : _ppm;
        if (PPM > _sf) throw Pothos::Exception("LoRaEncoder::work()", "failed check: PPM <= SF");
        lorab200_ctx *ctx = _ctx.get("LoRaEncoderB200", _sf);

        //every payload that is waiting, as one batch (payloads longer than 255 bytes do not fit the
        //one-byte header length field: the reference writes them modulo 256; reject them)
        std::vector<Pothos::Packet> pkts;
        size_t stride = 1;
        while (inPort->hasMessage())
        {
            auto msg = inPort->popMessage();
            pkts.push_back(msg.extract<Pothos::Packet>());
            stride = std::max(stride, pkts.back().payload.length);
        }
        if (stride > 255) throw Pothos::Exception("LoRaEncoderB200::work()", "payload longer than 255 bytes");
        const size_t n = pkts.size();
        lorab200_encoder_cfg cfg;
        cfg.ppm = int32_t(_ppm); cfg.rdd = int32_t(_rdd); cfg.explicit_header = _explicit; cfg.crc = _crc; cfg.whitening = _whitening;
        const size_t symStride = std::max<size_t>(8, lorab200_encode_symbols(int(_sf), &cfg, stride));
        std::vector<uint8_t> payloads(n*stride, 0);
        std::vector<int32_t> lengths(n), nSyms(n);
        for (size_t i = 0; i < n; i++)
        {
            lengths[i] = int32_t(pkts[i].payload.length);
            std::memcpy(payloads.data() + i*stride, pkts[i].payload.as<const void *>(), pkts[i].payload.length);
        }
        std::vector<uint16_t> symbols(n*symStride);
        check("LoRaEncoderB200::work()", ctx, lorab200_encode_host(ctx, payloads.data(), n, stride, lengths.data(), &cfg,
            symbols.data(), symStride, nSyms.data()));
        for (size_t i = 0; i < n; i++)
        {
            if (nSyms[i] < 0) throw Pothos::Exception("LoRaEncoderB200::work()", "symbol buffer too small");
            Pothos::Packet out;
            out.payload = Pothos::BufferChunk(typeid(uint16_t), size_t(nSyms[i]));
            std::memcpy(out.payload.as<void *>(), symbols.data() + i*symStride, out.payload.length);
            outPort->postMessage(out);
        }
    }

private:
    CtxHolder _ctx;
    size_t _sf, _ppm, _rdd;
    bool _explicit, _crc, _whitening;
};

static Pothos::BlockRegistry registerLoRaEncoderB200(LORAB200_ENCODER_PATH, &LoRaEncoderB200::make);

/***********************************************************************
 * |PothosDoc LoRa Decoder (B200)
 *
 * Decode LoRa modulation symbols into output bytes.  Drop-in replacement for the LoRa Decoder
 * block: all symbol packets waiting on the input port are decoded by one kernel launch on a B200.
 *
 * |category /LoRa
 * |keywords lora
 *
 * |param sf[Spread factor] |default 10
 * |param ppm[Symbol size] |default 0
 * |param cr[Coding Rate] |default "4/8"
 * |param explicit |default true
 * |param hdr |default false
 * |param dataLength[Data length] Implicit-mode payload length. |default 8
 * |param crcc |default false
 * |param whitening |default true
 * |param interleaving |default true
 * |param errorCheck |default false
 *
 * |factory /lora/lora_decoder()
 * |setter setSpreadFactor(sf)
 * |setter setSymbolSize(ppm)
 * |setter setCodingRate(cr)
 * |setter enableWhitening(whitening)
 * |setter enableCrcc(crcc)
 * |setter enableInterleaving(interleaving)
 * |setter enableExplicit(explicit)
 * |setter enableHdr(hdr)
 * |setter setDataLength(dataLength)
 * |setter enableErrorCheck(errorCheck)
 **********************************************************************/
class LoRaDecoderB200 : public Pothos::Block
{
public:
    LoRaDecoderB200(void): _sf(10), _ppm(0), _rdd(4), _whitening(true), _crcc(false), _interleaving(true),
        _errorCheck(false), _explicit(true), _hdr(false), _dataLength(8), _dropped(0)
    {
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDecoderB200, setSpreadFactor));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDecoderB200, setSymbolSize));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDecoderB200, setCodingRate));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDecoderB200, enableWhitening));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDecoderB200, enableCrcc));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDecoderB200, enableInterleaving));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDecoderB200, enableExplicit));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDecoderB200, enableHdr));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDecoderB200, setDataLength));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDecoderB200, enableErrorCheck));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaDecoderB200, getDropped));
        this->registerSignal("dropped");
        this->setupInput("0");
        this->setupOutput("0");
    }

    static Block *make(void) { return new LoRaDecoderB200(); }

    void setSpreadFactor(const size_t sf) { _sf = sf; }
    void setSymbolSize(const size_t ppm) { _ppm = ppm; }
    void setCodingRate(const std::string &cr) { _rdd = size_t(codingRate("LoRaDecoder", cr)); }
    void enableWhitening(const bool whitening) { _whitening = whitening; } //(not consulted by the reference decoder either)
    void enableInterleaving(const bool interleaving) { _interleaving = interleaving; }
    void enableExplicit(const bool e) { _explicit = e; }
    void enableHdr(const bool hdr) { _hdr = hdr; }
    void enableErrorCheck(const bool errorCheck) { _errorCheck = errorCheck; }
    void enableCrcc(const bool crcc) { _crcc = crcc; }
    void setDataLength(const size_t dataLength) { _dataLength = dataLength; }
    unsigned long long getDropped(void) const { return _dropped; }

    void activate(void)
    {
        _dropped = 0;
        this->emitSignal("dropped", _dropped);
    }

    void work(void)
    {
        auto inPort = this->input(0);
        auto outPort = this->output(0);
        if (not inPort->hasMessage()) return;
        const size_t PPM = (_ppm == 0) ? _sf : _ppm;
        if (PPM > _sf) throw Pothos::Exception("LoRaDecoder::work()", "failed check: PPM <= SF");
        lorab200_ctx *ctx = _ctx.get("LoRaDecoderB200", _sf);

        std::vector<Pothos::Packet> pkts;
        size_t stride = 8;
        while (inPort->hasMessage())
        {
            auto msg = inPort->popMessage();
            pkts.push_back(msg.extract<Pothos::Packet>());
            stride = std::max(stride, pkts.back().payload.elements());
        }
        const size_t n = pkts.size();
        if (not _interleaving)
        {
            //the reference's non-interleaved branch is a stub that produces nothing (LoRaDecoder.cpp:262-264)
            return;
        }
        std::vector<uint16_t> symbols(n*stride, 0);
        std::vector<int32_t> nSyms(n), lengths(n);
        for (size_t i = 0; i < n; i++)
        {
            nSyms[i] = int32_t(pkts[i].payload.elements());
            std::memcpy(symbols.data() + i*stride, pkts[i].payload.as<const void *>(), pkts[i].payload.length);
        }
        lorab200_decoder_cfg cfg;
        cfg.ppm = int32_t(_ppm); cfg.rdd = int32_t(_rdd); cfg.explicit_header = _explicit; cfg.hdr = _hdr;
        cfg.data_length = int32_t(_dataLength); cfg.crcc = _crcc; cfg.error_check = _errorCheck;
        const size_t byteStride = stride*2 + 16; //more than any packet of `stride` symbols can carry
        std::vector<uint8_t> bytes(n*byteStride);
        check("LoRaDecoderB200::work()", ctx, lorab200_decode_host(ctx, symbols.data(), n, stride, nSyms.data(), &cfg,
            bytes.data(), byteStride, lengths.data()));
        for (size_t i = 0; i < n; i++)
        {
            if (lengths[i] == LORAB200_DECODE_DROPPED) { this->drop(); continue; }
            if (lengths[i] < 0) continue; //fewer than 8 symbols: the reference returns without output or drop
            Pothos::Packet out;
            out.payload = Pothos::BufferChunk(typeid(uint8_t), size_t(lengths[i]));
            std::memcpy(out.payload.as<void *>(), bytes.data() + i*byteStride, out.payload.length);
            outPort->postMessage(out);
        }
    }

private:
    void drop(void)
    {
        _dropped++;
        this->emitSignal("dropped", _dropped);
    }

    CtxHolder _ctx;
    size_t _sf, _ppm, _rdd;
    bool _whitening, _crcc, _interleaving, _errorCheck, _explicit, _hdr;
    size_t _dataLength;
    unsigned long long _dropped;
};

static Pothos::BlockRegistry registerLoRaDecoderB200(LORAB200_DECODER_PATH, &LoRaDecoderB200::make);

/***********************************************************************
 * |PothosDoc LoRa Mod (B200)
 *
 * Modulate LoRa packets from symbols into a complex sample stream.  Drop-in replacement for the
 * LoRa Mod block: a whole packet (preamble, sync word, down-chirps, symbols, padding) is
 * synthesised by the GPU and streamed out of the block's output port.
 *
 * |category /LoRa
 * |keywords lora
 *
 * |param sf[Spread factor] Each symbol occupies 2^SF samples.
 * |default 10
 * |param sync[Sync word] |default 0x12
 * |param padding[Padding] Zero symbols after each packet. |default 1
 * |param ampl[Amplitude] |default 0.3
 * |param ovs[Oversampling] |default 1
 *
 * |factory /lora/lora_mod(sf)
 * |setter setSync(sync)
 * |setter setPadding(padding)
 * |setter setAmplitude(ampl)
 * |setter setOvs(ovs)
 **********************************************************************/
class LoRaModB200 : public Pothos::Block
{
public:
    LoRaModB200(const size_t sf): _sf(sf), _sync(0x12), _padding(1), _ampl(0.3f), _ovs(1), _ideal(false), _sent(0)
    {
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaModB200, setSync));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaModB200, setPadding));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaModB200, setAmplitude));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaModB200, setOvs));
        this->registerCall(this, POTHOS_FCN_TUPLE(LoRaModB200, setIdeal));
        this->setupInput(0);
        this->setupOutput(0, typeid(std::complex<float>));
        (void)_ctx.get("LoRaModB200", _sf);
    }

    static Block *make(const size_t sf) { return new LoRaModB200(sf); }

    void setSync(const unsigned char sync) { _sync = sync; }
    void setPadding(const size_t padding) { _padding = padding; }
    void setAmplitude(const float ampl) { _ampl = ampl; }
    void setOvs(const size_t ovs)
    {
        if (ovs < 1 or ovs > 256)
            throw Pothos::InvalidArgumentException("LoRaMod::setOvs(" + std::to_string(ovs) + ")", "invalid oversampling ratio");
        _ovs = ovs;
    }
    //! closed-form chirps (LORAB200_MOD_IDEAL, ovs 1 only) instead of the reference's recurrences
    void setIdeal(const bool ideal) { _ideal = ideal; }

    void activate(void)
    {
        _samples.clear();
        _sent = 0;
    }

    void work(void)
    {
        auto outPort = this->output(0);
        if (_sent == _samples.size())
        {
            //between packets: take the next one and synthesise all of its samples in one launch
            if (not this->input(0)->hasMessage()) return;
            auto msg = this->input(0)->popMessage();
            auto pkt = msg.extract<Pothos::Packet>();
            const size_t nSyms = pkt.payload.elements();
            lorab200_ctx *ctx = _ctx.get("LoRaModB200", _sf);
            const size_t total = lorab200_mod_samples(int(_sf), int(_ovs), nSyms, _padding);
            _samples.resize(total);
            _sent = 0;
            check("LoRaModB200::work()", ctx, lorab200_mod_host(ctx, pkt.payload.as<const uint16_t *>(), 1, nSyms, _sync,
                _padding, _ampl, int(_ovs), (_ideal and _ovs == 1) ? LORAB200_MOD_IDEAL : LORAB200_MOD_REFERENCE,
                reinterpret_cast<float *>(_samples.data())));
        }
        //stream the packet out as the output buffer allows
        const size_t room = outPort->elements();
        const size_t n = std::min(room, _samples.size() - _sent);
        if (n == 0) return;
        std::memcpy(outPort->buffer().as<std::complex<float> *>(), _samples.data() + _sent, n*sizeof(std::complex<float>));
        outPort->produce(n);
        _sent += n;
    }

private:
    CtxHolder _ctx;
    size_t _sf;
    unsigned char _sync;
    size_t _padding;
    float _ampl;
    size_t _ovs;
    bool _ideal;
    std::vector<std::complex<float>> _samples;
    size_t _sent;
};

static Pothos::BlockRegistry registerLoRaModB200(LORAB200_MOD_PATH, &LoRaModB200::make);
