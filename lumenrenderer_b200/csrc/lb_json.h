// lb_json.h — a small JSON DOM for the glTF ingest (host only). The reference parses glTF with fx-gltf + nlohmann::json
// (Lumen/vendor, not part of the path); only the document model glTF 2.0 needs is implemented here.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace lb {
namespace json {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false; double num = 0.0; std::string str;
    std::vector<Value> arr; std::vector<std::pair<std::string, Value>> obj;     // insertion order kept (attribute order matters to nobody, but it is cheap)

    bool is_null() const { return kind == Null; }
    bool is_object() const { return kind == Object; }
    bool is_array() const { return kind == Array; }
    const Value* find(const char* key) const {
        if (kind != Object) return nullptr;
        for (auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const char* key) const { return find(key) != nullptr; }
    const Value& operator[](const char* key) const { static const Value none; const Value* v = find(key); return v ? *v : none; }
    const Value& operator[](size_t i) const { static const Value none; return (kind == Array && i < arr.size()) ? arr[i] : none; }
    size_t size() const { return kind == Array ? arr.size() : (kind == Object ? obj.size() : 0); }
    double number(double dflt) const { return kind == Number ? num : dflt; }
    // clamped to +-2^62: the double -> int64 cast is undefined outside the int64 range (1e30, NaN), and callers only ever compare the
    // result with sizes and counts
    int64_t integer(int64_t dflt) const {
        if (kind != Number || num != num) return dflt;
        const double lim = 4611686018427387904.0;
        return num >= lim ? (int64_t)1 << 62 : (num <= -lim ? -((int64_t)1 << 62) : (int64_t)num);
    }
    const std::string& string() const { return str; }
};

class Parser {
public:
    Parser(const char* p, size_t n) : p_(p), end_(p + n) {}
    Value parse() { Value v = value(0); ws(); if (p_ != end_) err("trailing characters"); return v; }
private:
    const char* p_; const char* end_;
    [[noreturn]] void err(const char* what) const { throw std::runtime_error(std::string("JSON: ") + what); }
    void ws() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_; }
    bool lit(const char* s) { const size_t n = strlen(s); if ((size_t)(end_ - p_) >= n && memcmp(p_, s, n) == 0) { p_ += n; return true; } return false; }
    static void utf8(std::string& o, uint32_t c) {
        if (c < 0x80) o += (char)c;
        else if (c < 0x800) { o += (char)(0xC0 | (c >> 6)); o += (char)(0x80 | (c & 63)); }
        else if (c < 0x10000) { o += (char)(0xE0 | (c >> 12)); o += (char)(0x80 | ((c >> 6) & 63)); o += (char)(0x80 | (c & 63)); }
        else { o += (char)(0xF0 | (c >> 18)); o += (char)(0x80 | ((c >> 12) & 63)); o += (char)(0x80 | ((c >> 6) & 63)); o += (char)(0x80 | (c & 63)); }
    }
    uint32_t hex4() {
        if (end_ - p_ < 4) err("bad \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; ++i) { const char c = *p_++; v <<= 4; if (c >= '0' && c <= '9') v |= c - '0'; else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10; else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10; else err("bad \\u escape"); }
        return v;
    }
    std::string string() {
        if (p_ >= end_ || *p_ != '"') err("expected string");
        ++p_; std::string o;
        while (p_ < end_ && *p_ != '"') {
            char c = *p_++;
            if (c != '\\') { o += c; continue; }
            if (p_ >= end_) err("bad escape");
            c = *p_++;
            switch (c) {
                case '"': o += '"'; break; case '\\': o += '\\'; break; case '/': o += '/'; break;
                case 'b': o += '\b'; break; case 'f': o += '\f'; break; case 'n': o += '\n'; break; case 'r': o += '\r'; break; case 't': o += '\t'; break;
                case 'u': { uint32_t cp = hex4(); if (cp >= 0xD800 && cp < 0xDC00 && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') { p_ += 2; const uint32_t lo = hex4(); cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); } utf8(o, cp); break; }
                default: err("bad escape");
            }
        }
        if (p_ >= end_) err("unterminated string");
        ++p_;
        return o;
    }
    Value value(int depth) {
        if (depth > 256) err("nesting too deep");
        ws();
        if (p_ >= end_) err("unexpected end");
        Value v;
        const char c = *p_;
        if (c == '{') {
            ++p_; v.kind = Value::Object; ws();
            if (p_ < end_ && *p_ == '}') { ++p_; return v; }
            for (;;) {
                ws(); std::string k = string(); ws();
                if (p_ >= end_ || *p_ != ':') err("expected ':'");
                ++p_; v.obj.emplace_back(std::move(k), value(depth + 1)); ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == '}') { ++p_; return v; }
                err("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++p_; v.kind = Value::Array; ws();
            if (p_ < end_ && *p_ == ']') { ++p_; return v; }
            for (;;) {
                v.arr.push_back(value(depth + 1)); ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == ']') { ++p_; return v; }
                err("expected ',' or ']'");
            }
        }
        if (c == '"') { v.kind = Value::String; v.str = string(); return v; }
        if (lit("true")) { v.kind = Value::Bool; v.b = true; return v; }
        if (lit("false")) { v.kind = Value::Bool; v.b = false; return v; }
        if (lit("null")) return v;
        // number
        const char* s = p_;
        if (p_ < end_ && (*p_ == '-' || *p_ == '+')) ++p_;
        while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '-' || *p_ == '+')) ++p_;
        if (p_ == s) err("unexpected character");
        const std::string tok(s, p_);
        char* e = nullptr; v.num = strtod(tok.c_str(), &e);
        if (e != tok.c_str() + tok.size()) err("bad number");
        // an INTEGER token is an integer in nlohmann-json (which the reference's fx-gltf reads with): "-0" is the integer 0 and becomes +0.0f,
        // "-0.0" keeps its sign. Exporters write both (the reference's skycastle asset holds 1 659 "-0" matrix elements).
        if (v.num == 0.0 && tok.find_first_of(".eE") == std::string::npos) v.num = 0.0;
        v.kind = Value::Number;
        return v;
    }
};

inline Value parse(const char* p, size_t n) { return Parser(p, n).parse(); }

} // namespace json
} // namespace lb
