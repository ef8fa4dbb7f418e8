// Minimal JSON reader for Piper voice configs (`*.onnx.json`; schema = ModelConfig at
// piper/src/lib.rs:112-158, parsed there by serde_json).  Objects, arrays, strings (with \uXXXX
// and surrogate pairs -> UTF-8), numbers, true/false/null.  Throws std::runtime_error.
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include <cstdlib>

namespace sbjson {

struct Value;
using ValuePtr = std::shared_ptr<Value>;

struct Value {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<ValuePtr> arr;
    std::vector<std::pair<std::string, ValuePtr>> obj;   // insertion order kept

    const Value* get(const std::string& k) const {
        for (auto& kv : obj) if (kv.first == k) return kv.second.get();
        return nullptr;
    }
    bool is_null() const { return kind == Null; }
};

class Parser {
public:
    explicit Parser(const std::string& s) : s_(s) {}
    ValuePtr parse() {
        ValuePtr v = value();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& s_;
    size_t p_ = 0;

    [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("JSON: ") + m + " at byte " + std::to_string(p_)); }
    void ws() { while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r')) p_++; }
    char peek() { ws(); if (p_ >= s_.size()) fail("unexpected end"); return s_[p_]; }

    static void put_utf8(std::string& o, unsigned cp) {
        if (cp < 0x80) o += (char)cp;
        else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
        else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
    }
    unsigned hex4() {
        if (p_ + 4 > s_.size()) fail("bad \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; i++) {
            char c = s_[p_++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else fail("bad hex digit");
        }
        return v;
    }
    std::string string() {
        if (peek() != '"') fail("expected string");
        p_++;
        std::string o;
        while (true) {
            if (p_ >= s_.size()) fail("unterminated string");
            char c = s_[p_++];
            if (c == '"') break;
            if (c == '\\') {
                if (p_ >= s_.size()) fail("bad escape");
                char e = s_[p_++];
                switch (e) {
                    case '"': o += '"'; break;
                    case '\\': o += '\\'; break;
                    case '/': o += '/'; break;
                    case 'b': o += '\b'; break;
                    case 'f': o += '\f'; break;
                    case 'n': o += '\n'; break;
                    case 'r': o += '\r'; break;
                    case 't': o += '\t'; break;
                    case 'u': {
                        unsigned cp = hex4();
                        if (cp >= 0xD800 && cp < 0xDC00 && p_ + 1 < s_.size() && s_[p_] == '\\' && s_[p_ + 1] == 'u') {
                            p_ += 2;
                            unsigned lo = hex4();
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        put_utf8(o, cp);
                        break;
                    }
                    default: fail("bad escape");
                }
            } else o += c;
        }
        return o;
    }
    ValuePtr value() {
        char c = peek();
        auto v = std::make_shared<Value>();
        if (c == '{') {
            p_++;
            v->kind = Value::Obj;
            if (peek() == '}') { p_++; return v; }
            while (true) {
                std::string k = string();
                if (peek() != ':') fail("expected ':'");
                p_++;
                v->obj.emplace_back(k, value());
                char d = peek();
                p_++;
                if (d == '}') break;
                if (d != ',') fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            p_++;
            v->kind = Value::Arr;
            if (peek() == ']') { p_++; return v; }
            while (true) {
                v->arr.push_back(value());
                char d = peek();
                p_++;
                if (d == ']') break;
                if (d != ',') fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v->kind = Value::Str;
            v->str = string();
        } else if (s_.compare(p_, 4, "true") == 0) { v->kind = Value::Bool; v->b = true; p_ += 4; }
        else if (s_.compare(p_, 5, "false") == 0) { v->kind = Value::Bool; v->b = false; p_ += 5; }
        else if (s_.compare(p_, 4, "null") == 0) { v->kind = Value::Null; p_ += 4; }
        else {
            const char* st = s_.c_str() + p_;
            char* en = nullptr;
            v->num = std::strtod(st, &en);
            if (en == st) fail("unexpected token");
            v->kind = Value::Num;
            p_ += (size_t)(en - st);
        }
        return v;
    }
};

inline ValuePtr parse(const std::string& s) { return Parser(s).parse(); }

}  // namespace sbjson
