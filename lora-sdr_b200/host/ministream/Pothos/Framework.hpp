// ministream -- a small, header-only stand-in for the part of the Pothos framework
// API that LoRa blocks use.  Pothos itself cannot be
// installed here (no network, no Poco), so this surface lets
//   * our drop-in block shells (lora-sdr_b200/host/blocks/*.cpp) compile and run, and
//   * the UNMODIFIED reference blocks (/root/reference/LoRa*.cpp) compile in place
//     into oracle/_ref for the loopback parity tests,
// against the same class and method names.  It is written from the API's documented
// behaviour, not from Pothos sources.  With a real Pothos install the block shells
// build against <Pothos/Framework.hpp> proper.
//
// Provided: Exception types, DType, BufferChunk, Object, Label, Packet,
// BufferManager(+Args), InputPort, OutputPort, Block (setupInput/Output, registerCall,
// registerSignal, emitSignal, call), BlockRegistry, POTHOS_FCN_TUPLE, and
// ministream::Topology -- a single-threaded scheduler with stream and message edges.
#pragma once
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <typeindex>
#include <typeinfo>
#include <vector>

#include "Config.hpp"

namespace Pothos {

// ---- exceptions ---------------------------------------------------------------------
class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string &msg) : std::runtime_error(msg) {}
    Exception(const std::string &msg, const std::string &arg) : std::runtime_error(msg + ": " + arg) {}
    std::string displayText() const { return what(); }
};
class InvalidArgumentException : public Exception {
public:
    using Exception::Exception;
};
class NotFoundException : public Exception {
public:
    using Exception::Exception;
};

// ---- DType ----------------------------------------------------------------------------
class DType {
public:
    DType() : _size(1), _name("") {}
    DType(const std::type_info &t) { init(std::type_index(t)); }
    DType(const std::string &name) : _size(1), _name(name) {}
    DType(const char *name) : _size(1), _name(name) {}
    size_t size() const { return _size; }
    const std::string &name() const { return _name; }
    bool operator==(const DType &o) const { return _name == o._name && _size == o._size; }

private:
    void init(std::type_index t)
    {
#define MS_DT(T, N) if (t == std::type_index(typeid(T))) { _size = sizeof(T); _name = N; return; }
        MS_DT(int8_t, "int8") MS_DT(uint8_t, "uint8") MS_DT(char, "int8")
        MS_DT(int16_t, "int16") MS_DT(uint16_t, "uint16")
        MS_DT(int32_t, "int32") MS_DT(uint32_t, "uint32")
        MS_DT(int64_t, "int64") MS_DT(uint64_t, "uint64")
        MS_DT(float, "float32") MS_DT(double, "float64")
        MS_DT(std::complex<float>, "complex_float32") MS_DT(std::complex<double>, "complex_float64")
        MS_DT(std::complex<int16_t>, "complex_int16") MS_DT(std::complex<int8_t>, "complex_int8")
#undef MS_DT
        _size = 1; _name = "custom";
    }
    size_t _size;
    std::string _name;
};

// ---- BufferChunk: a ref-counted view of bytes --------------------------------------------
class BufferChunk {
public:
    size_t address = 0; // writable by users, like the real class
    size_t length = 0;  // bytes
    DType dtype;

    BufferChunk() {}
    explicit BufferChunk(size_t numBytes) { alloc(numBytes); }
    BufferChunk(const DType &dt, size_t numElems) : dtype(dt) { alloc(numElems * dt.size()); }
    // non-owning view (used by the scheduler for port buffers)
    static BufferChunk view(void *p, size_t bytes, const DType &dt)
    {
        BufferChunk b; b.address = size_t(p); b.length = bytes; b.dtype = dt; return b;
    }
    template <typename T> T as() const { return reinterpret_cast<T>(address); }
    size_t elements() const { return length / dtype.size(); }
    explicit operator bool() const { return address != 0; }

private:
    void alloc(size_t bytes)
    {
        _store = std::shared_ptr<uint8_t>(new uint8_t[bytes ? bytes : 1](), std::default_delete<uint8_t[]>());
        address = size_t(_store.get());
        length = bytes;
    }
    std::shared_ptr<uint8_t> _store;
};

// ---- Object: holds any copyable value ---------------------------------------------------------
class Object {
public:
    Object() {}
    template <typename T, typename = typename std::enable_if<!std::is_same<typename std::decay<T>::type, Object>::value>::type>
    Object(T &&v) : _h(std::make_shared<Holder<typename std::decay<T>::type>>(std::forward<T>(v))) {}
    explicit operator bool() const { return bool(_h); }
    const std::type_info &type() const { return _h ? _h->type() : typeid(void); }
    template <typename T> const T &extract() const
    {
        auto p = dynamic_cast<const Holder<T> *>(_h.get());
        if (!p) throw Exception("Pothos::Object::extract", std::string("type mismatch, holds ") + (_h ? _h->type().name() : "null"));
        return p->v;
    }
    // numeric / exact conversion used by call() argument passing
    template <typename T> T convert() const
    {
        if (auto p = dynamic_cast<const Holder<T> *>(_h.get())) return p->v;
        return convertNum<T>(std::is_arithmetic<T>());
    }

private:
    struct Base {
        virtual ~Base() {}
        virtual const std::type_info &type() const = 0;
        virtual bool num(long double &out) const = 0;
    };
    template <typename T> struct Holder : Base {
        template <typename U> explicit Holder(U &&u) : v(std::forward<U>(u)) {}
        const std::type_info &type() const override { return typeid(T); }
        bool num(long double &out) const override { return toNum(out, std::is_arithmetic<T>()); }
        bool toNum(long double &out, std::true_type) const { out = static_cast<long double>(v); return true; }
        bool toNum(long double &, std::false_type) const { return false; }
        T v;
    };
    template <typename T> T convertNum(std::true_type) const
    {
        long double x;
        if (_h && _h->num(x)) return static_cast<T>(x);
        throw Exception("Pothos::Object::convert", "not convertible");
    }
    template <typename T> T convertNum(std::false_type) const { throw Exception("Pothos::Object::convert", "not convertible"); }
    std::shared_ptr<Base> _h;
};

// ---- Label / Packet -----------------------------------------------------------------------
class Label {
public:
    std::string id;
    Object data;
    unsigned long long index = 0;
    size_t width = 1;
    Label() {}
    Label(const std::string &id_, const Object &data_, unsigned long long index_, size_t width_ = 1)
        : id(id_), data(data_), index(index_), width(width_) {}
};

class Packet {
public:
    BufferChunk payload;
    std::map<std::string, Object> metadata;
    std::vector<Label> labels;
};

// ---- BufferManager (sizes only; the scheduler owns the memory) ---------------------------------
struct BufferManagerArgs {
    size_t numBuffers = 4;
    size_t bufferSize = 8 * 1024;
    long nodeAffinity = -1;
};
class BufferManager {
public:
    typedef std::shared_ptr<BufferManager> Sptr;
    virtual ~BufferManager() {}
    static Sptr make(const std::string &name, const BufferManagerArgs &args)
    {
        auto p = std::make_shared<BufferManager>();
        p->name = name; p->args = args;
        return p;
    }
    // A custom manager (a block's getInputBufferManager / getOutputBufferManager) owns where the
    // port's bytes live: the scheduler takes the port's slab from allocate() and gives it back
    // with release().  The default is the heap.  (Real Pothos: a BufferManager subclass hands out
    // ManagedBuffers over its own SharedBuffer slabs; same ownership, richer interface.)
    virtual void *allocate(size_t bytes) { return std::malloc(bytes); }
    virtual void release(void *p) { std::free(p); }
    virtual bool custom() const { return false; }
    std::string name;
    BufferManagerArgs args;
};

class Block;

// ---- ports ----------------------------------------------------------------------------------
class InputPort {
public:
    ~InputPort() { if (_q) _mgr->release(_q); }
    size_t elements() const { return (_size - _head) / _dtype.size(); }
    BufferChunk buffer() { return BufferChunk::view(_q + _head, _size - _head, _dtype); }
    void consume(size_t n) { _consumed += n; }
    bool hasMessage() const { return _mhead < _msgs.size(); }
    Object popMessage()
    {
        if (!hasMessage()) return Object();
        _progress = true;
        return _msgs[_mhead++];
    }
    void setReserve(size_t n) { _reserve = n; }
    const std::string &name() const { return _name; }
    const DType &dtype() const { return _dtype; }
    unsigned long long totalElements() const { return _total; }

    // scheduler side: the port's bytes live in ONE slab from the port's buffer manager (heap by
    // default, e.g. page-locked memory when the consuming block says so); the unread bytes are
    // moved to the front when the slab would overflow, the slab grows geometrically
    void _setManager(const BufferManager::Sptr &m)
    {
        if (!m || !m->custom()) return;
        if (_q) {
            uint8_t *n = static_cast<uint8_t *>(m->allocate(_cap));
            if (!n) throw Exception("InputPort", "buffer manager: out of memory");
            std::memcpy(n, _q + _head, _size - _head);
            _mgr->release(_q);
            _q = n; _size -= _head; _head = 0;
        }
        _mgr = m;
    }
    void _pushBytes(const void *p, size_t bytes)
    {
        if (_size + bytes > _cap) {
            const size_t live = _size - _head;
            if (_head >= live + bytes && live + bytes <= _cap) {
                std::memmove(_q, _q + _head, live);            // enough room once the consumed front is dropped
            } else {
                size_t ncap = _cap ? _cap : size_t(1) << 16;
                while (ncap < 2 * (live + bytes)) ncap *= 2;
                uint8_t *n = static_cast<uint8_t *>(_mgr->allocate(ncap));
                if (!n) throw Exception("InputPort", "buffer manager: out of memory");
                if (live) std::memcpy(n, _q + _head, live);
                if (_q) _mgr->release(_q);
                _q = n; _cap = ncap;
            }
            _size = live; _head = 0;
        }
        std::memcpy(_q + _size, p, bytes);
        _size += bytes;
    }
    // a framework sizes its port slabs when the topology is committed, not in the data path: harnesses
    // that push a whole stream at once reserve its size first (page-locked slabs are slow to allocate)
    void _reserveBytes(size_t cap)
    {
        if (cap <= _cap) return;
        uint8_t *n = static_cast<uint8_t *>(_mgr->allocate(cap));
        if (!n) throw Exception("InputPort", "buffer manager: out of memory");
        const size_t live = _size - _head;
        if (live) std::memcpy(n, _q + _head, live);
        if (_q) _mgr->release(_q);
        _q = n; _cap = cap; _size = live; _head = 0;
    }
    void _pushMessage(const Object &o) { _msgs.push_back(o); }
    bool _commit()
    {
        bool did = _consumed > 0 || _progress;
        _head += _consumed * _dtype.size();
        if (_head > _size) _head = _size;
        _total += _consumed;
        _consumed = 0; _progress = false;
        if (_mhead > 1024 && _mhead == _msgs.size()) { _msgs.clear(); _mhead = 0; }
        return did;
    }
    size_t _reserveElems() const { return _reserve; }

    std::string _name;
    DType _dtype;

private:
    BufferManager::Sptr _mgr = std::make_shared<BufferManager>();
    uint8_t *_q = nullptr;
    size_t _cap = 0, _size = 0;
    size_t _head = 0, _consumed = 0, _reserve = 0, _mhead = 0;
    unsigned long long _total = 0;
    bool _progress = false;
    std::vector<Object> _msgs;
};

class OutputPort {
public:
    BufferChunk buffer()
    {
        if (_scratch.size() < _bufferBytes) _scratch.resize(_bufferBytes);
        return BufferChunk::view(_scratch.data(), _scratch.size(), _dtype);
    }
    size_t elements() const { return _bufferBytes / _dtype.size(); }
    void produce(size_t n) { _produced += n; }
    template <typename T> void postMessage(T &&m) { _outMsgs.emplace_back(Object(std::forward<T>(m))); }
    void postLabel(const Label &l) { _outLabels.push_back(l); }
    void setReserve(size_t n) { _reserve = n; }
    const std::string &name() const { return _name; }
    const DType &dtype() const { return _dtype; }
    unsigned long long totalElements() const { return _total; }

    std::string _name;
    DType _dtype;
    size_t _bufferBytes = 64 * 1024;
    std::vector<uint8_t> _scratch;
    size_t _produced = 0, _reserve = 0;
    unsigned long long _total = 0;
    std::vector<Object> _outMsgs;      // posted during the current work()
    std::vector<Label> _outLabels;     // posted during the current work() (index relative to the buffer)
    // sinks for unconnected ports / inspection (filled by the scheduler)
    std::vector<Object> collectedMessages;
    std::vector<Label> collectedLabels; // index made absolute (stream position)
    std::vector<uint8_t> collectedBytes;
    bool collect = false;
    std::vector<std::pair<Block *, std::string>> _subs;
};

// ---- Block ------------------------------------------------------------------------------------
class Block {
public:
    virtual ~Block() {}
    virtual void work() {}
    virtual void activate() {}
    virtual void deactivate() {}
    virtual BufferManager::Sptr getInputBufferManager(const std::string &, const std::string &) { return BufferManager::Sptr(); }
    virtual BufferManager::Sptr getOutputBufferManager(const std::string &, const std::string &) { return BufferManager::Sptr(); }

    void setupInput(size_t i, const DType &dt = DType(), const std::string & = "") { setupInput(std::to_string(i), dt); }
    void setupInput(const std::string &name, const DType &dt = DType(), const std::string & = "")
    {
        auto p = std::make_shared<InputPort>(); p->_name = name; p->_dtype = dt; _in[name] = p; _inOrder.push_back(name);
    }
    void setupOutput(size_t i, const DType &dt = DType(), const std::string & = "") { setupOutput(std::to_string(i), dt); }
    void setupOutput(const std::string &name, const DType &dt = DType(), const std::string & = "")
    {
        auto p = std::make_shared<OutputPort>(); p->_name = name; p->_dtype = dt; _out[name] = p; _outOrder.push_back(name);
    }
    InputPort *input(size_t i) { return input(std::to_string(i)); }
    InputPort *input(const std::string &n) { auto it = _in.find(n); if (it == _in.end()) throw NotFoundException("Block::input", n); return it->second.get(); }
    OutputPort *output(size_t i) { return output(std::to_string(i)); }
    OutputPort *output(const std::string &n) { auto it = _out.find(n); if (it == _out.end()) throw NotFoundException("Block::output", n); return it->second.get(); }
    const std::vector<std::string> &inputNames() const { return _inOrder; }
    const std::vector<std::string> &outputNames() const { return _outOrder; }

    // registerCall(this, "name", &Class::method) for any arity, const or not
    template <typename C, typename R, typename... A>
    void registerCall(C *self, const std::string &name, R (C::*m)(A...)) { bind<C, R, A...>(self, name, [self, m](A... a) -> R { return (self->*m)(a...); }); }
    template <typename C, typename R, typename... A>
    void registerCall(C *self, const std::string &name, R (C::*m)(A...) const) { bind<C, R, A...>(self, name, [self, m](A... a) -> R { return (self->*m)(a...); }); }

    void registerSignal(const std::string &name) { _signals[name]; }
    template <typename... A> void emitSignal(const std::string &name, A &&...a)
    {
        std::vector<Object> args{Object(std::forward<A>(a))...};
        signalLog.emplace_back(name, args);
        auto it = _signals.find(name);
        if (it != _signals.end()) for (auto &slot : it->second) slot(args);
    }
    void connectSignal(const std::string &name, std::function<void(const std::vector<Object> &)> slot) { _signals[name].push_back(std::move(slot)); }

    // call a registered method by name (what Pothos::Proxy::call does for setters)
    template <typename... A> Object call(const std::string &name, A &&...a)
    {
        auto it = _calls.find(name);
        if (it == _calls.end()) throw NotFoundException("Block::call", name);
        return it->second(std::vector<Object>{Object(std::forward<A>(a))...});
    }
    template <typename R, typename... A> R call(const std::string &name, A &&...a) { return call(name, std::forward<A>(a)...).template convert<R>(); }

    std::vector<std::pair<std::string, std::vector<Object>>> signalLog;
    std::string blockName;

private:
    template <typename T> struct Arg { typedef typename std::decay<T>::type type; };
    template <typename C, typename R, typename... A, typename F>
    void bind(C *, const std::string &name, F f)
    {
        _calls[name] = [f, name](const std::vector<Object> &args) -> Object {
            if (args.size() != sizeof...(A)) throw InvalidArgumentException("Block::call(" + name + ")", "wrong argument count");
            return invoke<R, A...>(f, args, std::index_sequence_for<A...>());
        };
    }
    template <typename R, typename... A, typename F, size_t... I>
    static Object invoke(F &f, const std::vector<Object> &args, std::index_sequence<I...>)
    {
        return ret<R>(f, args[I].template convert<typename Arg<A>::type>()...);
    }
    template <typename R, typename F, typename... V>
    static typename std::enable_if<std::is_void<R>::value, Object>::type ret(F &f, V &&...v) { f(std::forward<V>(v)...); return Object(); }
    template <typename R, typename F, typename... V>
    static typename std::enable_if<!std::is_void<R>::value, Object>::type ret(F &f, V &&...v) { return Object(f(std::forward<V>(v)...)); }

    std::map<std::string, std::shared_ptr<InputPort>> _in;
    std::map<std::string, std::shared_ptr<OutputPort>> _out;
    std::vector<std::string> _inOrder, _outOrder;
    std::map<std::string, std::function<Object(const std::vector<Object> &)>> _calls;
    std::map<std::string, std::vector<std::function<void(const std::vector<Object> &)>>> _signals;
};

// ---- BlockRegistry --------------------------------------------------------------------------
class BlockRegistry {
public:
    typedef std::function<Block *(const std::vector<Object> &)> Factory;
    BlockRegistry(const std::string &path, Block *(*f)(void)) { table()[path] = [f](const std::vector<Object> &) { return f(); }; }
    template <typename A0> BlockRegistry(const std::string &path, Block *(*f)(A0))
    {
        table()[path] = [f, path](const std::vector<Object> &a) {
            if (a.size() != 1) throw InvalidArgumentException("BlockRegistry::make(" + path + ")", "expected 1 argument");
            return f(a[0].convert<typename std::decay<A0>::type>());
        };
    }
    template <typename A0, typename A1> BlockRegistry(const std::string &path, Block *(*f)(A0, A1))
    {
        table()[path] = [f, path](const std::vector<Object> &a) {
            if (a.size() != 2) throw InvalidArgumentException("BlockRegistry::make(" + path + ")", "expected 2 arguments");
            return f(a[0].convert<typename std::decay<A0>::type>(), a[1].convert<typename std::decay<A1>::type>());
        };
    }
    template <typename... A> static Block *make(const std::string &path, A &&...a)
    {
        auto it = table().find(path);
        if (it == table().end()) throw NotFoundException("BlockRegistry::make", path);
        Block *b = it->second(std::vector<Object>{Object(std::forward<A>(a))...});
        b->blockName = path;
        return b;
    }
    static bool doesBlockExist(const std::string &path) { return table().count(path) != 0; }
    static std::vector<std::string> paths() { std::vector<std::string> r; for (auto &kv : table()) r.push_back(kv.first); return r; }

private:
    static std::map<std::string, Factory> &table() { static std::map<std::string, Factory> t; return t; }
};

} // namespace Pothos

#define POTHOS_FCN_TUPLE(classPath, functionName) #functionName, &classPath::functionName

// ---- the scheduler ------------------------------------------------------------------------------
namespace ministream {

class Topology {
public:
    // stream or message edge: everything produced / posted on src.srcPort goes to dst.dstPort
    void connect(Pothos::Block *src, const std::string &srcPort, Pothos::Block *dst, const std::string &dstPort)
    {
        add(src); add(dst);
        src->output(srcPort)->_subs.emplace_back(dst, dstPort);
    }
    void add(Pothos::Block *b) { for (auto x : _blocks) if (x == b) return; _blocks.push_back(b); }

    // (re)size port buffers from the blocks' buffer managers and activate
    void commit()
    {
        for (auto b : _blocks) {
            for (auto &n : b->outputNames()) {
                auto mgr = b->getOutputBufferManager(n, "");
                if (mgr) b->output(n)->_bufferBytes = std::max(b->output(n)->_bufferBytes, mgr->args.bufferSize);
            }
            for (auto &n : b->inputNames()) b->input(n)->_setManager(b->getInputBufferManager(n, ""));
            b->activate();
        }
        _active = true;
    }

    // run every block until a whole round makes no progress (Topology::waitInactive)
    size_t run(size_t maxRounds = size_t(-1))
    {
        if (!_active) commit();
        size_t rounds = 0;
        bool progress = true;
        while (progress && rounds < maxRounds) {
            progress = false;
            for (auto b : _blocks) progress |= step(b);
            ++rounds;
        }
        return rounds;
    }

    static void feedStream(Pothos::Block *b, const std::string &port, const void *data, size_t bytes) { b->input(port)->_pushBytes(data, bytes); }
    static void feedMessage(Pothos::Block *b, const std::string &port, const Pothos::Object &o) { b->input(port)->_pushMessage(o); }

private:
    bool runnable(Pothos::Block *b)
    {
        if (b->inputNames().empty()) return true;
        for (auto &n : b->inputNames()) {
            auto p = b->input(n);
            if (p->hasMessage()) return true;
            const size_t need = p->_reserveElems() > 0 ? p->_reserveElems() : 1;
            if (p->elements() >= need) return true;
        }
        return _sourcesLike.count(b) != 0;
    }
    bool step(Pothos::Block *b)
    {
        // blocks with only message inputs and internal state (e.g. a modulator mid-packet)
        // must be polled: treat every block as runnable, detect progress afterwards
        (void)runnable(b);
        for (auto &n : b->outputNames()) { auto o = b->output(n); o->_produced = 0; o->_outMsgs.clear(); o->_outLabels.clear(); }
        b->work();
        bool did = false;
        for (auto &n : b->inputNames()) did |= b->input(n)->_commit();
        for (auto &n : b->outputNames()) {
            auto o = b->output(n);
            const size_t bytes = o->_produced * o->dtype().size();
            if (bytes || !o->_outMsgs.empty() || !o->_outLabels.empty()) did = true;
            for (auto &l : o->_outLabels) { Pothos::Label a = l; a.index += o->_total; if (o->collect || o->_subs.empty()) o->collectedLabels.push_back(a); }
            if (bytes) {
                if (o->collect || o->_subs.empty()) {
                    if (o->collect) o->collectedBytes.insert(o->collectedBytes.end(), o->_scratch.begin(), o->_scratch.begin() + long(bytes));
                }
                for (auto &s : o->_subs) s.first->input(s.second)->_pushBytes(o->_scratch.data(), bytes);
                o->_total += o->_produced;
            }
            for (auto &m : o->_outMsgs) {
                if (o->_subs.empty() || o->collect) o->collectedMessages.push_back(m);
                for (auto &s : o->_subs) s.first->input(s.second)->_pushMessage(m);
            }
            o->_produced = 0; o->_outMsgs.clear(); o->_outLabels.clear();
        }
        return did;
    }
    std::vector<Pothos::Block *> _blocks;
    std::map<Pothos::Block *, int> _sourcesLike;
    bool _active = false;
};

} // namespace ministream
