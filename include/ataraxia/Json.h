// Json.h — a minimal JSON value, reader and writer: just what scene.json needs (the reference
// uses nlohmann::json 3.11.3, Engine/src/Utils.cpp). Objects keep their keys sorted (std::map, as
// nlohmann's default object type does), numbers are doubles, dump(4) indents like nlohmann's.
// Missing keys and type mismatches throw (std::out_of_range / std::runtime_error), where nlohmann
// throws json::out_of_range / json::type_error: both escape Utils::importScene uncaught.
#pragma once
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace atx
{
class Json
{
public:
    enum class Type { Null, Bool, Number, Integer, String, Array, Object };
    using Array = std::vector<Json>;
    using Object = std::map<std::string, Json>;

    Json() = default;
    Json(bool b) : m_type(Type::Bool), m_bool(b) {}
    Json(double d) : m_type(Type::Number), m_num(d) {}
    Json(float f) : m_type(Type::Number), m_num(static_cast<double>(f)) {}
    Json(int i) : m_type(Type::Integer), m_num(i) {}
    Json(const char* s) : m_type(Type::String), m_str(s) {}
    Json(const std::string& s) : m_type(Type::String), m_str(s) {}
    static Json array() { Json j; j.m_type = Type::Array; return j; }
    static Json object() { Json j; j.m_type = Type::Object; return j; }

    Type type() const { return m_type; }
    bool isNull() const { return m_type == Type::Null; }
    bool contains(const std::string& key) const { return m_type == Type::Object && m_obj.count(key) != 0; }
    size_t size() const { return m_type == Type::Array ? m_arr.size() : (m_type == Type::Object ? m_obj.size() : 0); }

    // object access: creates on a mutable value, throws on a const one
    Json& operator[](const std::string& key)
    {
        if (m_type == Type::Null) m_type = Type::Object;
        if (m_type != Type::Object) throw std::runtime_error("json: not an object");
        return m_obj[key];
    }
    const Json& at(const std::string& key) const
    {
        if (m_type != Type::Object) throw std::runtime_error("json: not an object (key '" + key + "')");
        auto it = m_obj.find(key);
        if (it == m_obj.end()) throw std::out_of_range("json: key '" + key + "' not found");
        return it->second;
    }
    const Json& operator[](const std::string& key) const { return at(key); }
    const Json& at(size_t i) const
    {
        if (m_type != Type::Array) throw std::runtime_error("json: not an array");
        if (i >= m_arr.size()) throw std::out_of_range("json: array index out of range");
        return m_arr[i];
    }
    const Json& operator[](size_t i) const { return at(i); }
    const Json& operator[](int i) const { return at(static_cast<size_t>(i)); }
    void push_back(const Json& v)
    {
        if (m_type == Type::Null) m_type = Type::Array;
        if (m_type != Type::Array) throw std::runtime_error("json: not an array");
        m_arr.push_back(v);
    }
    const Array& items() const
    {
        static const Array empty;
        return m_type == Type::Array ? m_arr : empty;
    }

    // value access (nlohmann's get<T>(): numbers convert between arithmetic types)
    double number() const
    {
        if (m_type != Type::Number && m_type != Type::Integer) throw std::runtime_error("json: type must be number");
        return m_num;
    }
    float getFloat() const { return static_cast<float>(number()); } // double -> float narrowing, as get<float>()
    int getInt() const { return static_cast<int>(number()); }
    bool getBool() const
    {
        if (m_type != Type::Bool) throw std::runtime_error("json: type must be boolean");
        return m_bool;
    }
    const std::string& getString() const
    {
        if (m_type != Type::String) throw std::runtime_error("json: type must be string");
        return m_str;
    }

    // ---- reader ----
    static Json parse(const std::string& text)
    {
        size_t pos = 0;
        Json v = parseValue(text, pos);
        skipWs(text, pos);
        if (pos != text.size()) throw std::runtime_error("json: trailing characters");
        return v;
    }

    // ---- writer (nlohmann dump(indent)) ----
    std::string dump(int indent = -1) const
    {
        std::string out;
        write(out, indent, 0);
        return out;
    }

private:
    Type m_type = Type::Null;
    bool m_bool = false;
    double m_num = 0.0;
    std::string m_str;
    Array m_arr;
    Object m_obj;

    static void skipWs(const std::string& t, size_t& p)
    {
        while (p < t.size() && (t[p] == ' ' || t[p] == '\t' || t[p] == '\n' || t[p] == '\r')) p++;
    }
    static constexpr int kMaxDepth = 256; // nesting cap: a hostile file cannot exhaust the stack
    static Json parseValue(const std::string& t, size_t& p, int depth = 0)
    {
        if (depth > kMaxDepth) throw std::runtime_error("json: nesting too deep");
        skipWs(t, p);
        if (p >= t.size()) throw std::runtime_error("json: unexpected end of input");
        const char c = t[p];
        if (c == '{')
        {
            Json j = object();
            p++;
            skipWs(t, p);
            if (p < t.size() && t[p] == '}') { p++; return j; }
            while (true)
            {
                skipWs(t, p);
                if (p >= t.size() || t[p] != '"') throw std::runtime_error("json: expected a string key");
                const std::string key = parseString(t, p);
                skipWs(t, p);
                if (p >= t.size() || t[p] != ':') throw std::runtime_error("json: expected ':'");
                p++;
                j.m_obj[key] = parseValue(t, p, depth + 1);
                skipWs(t, p);
                if (p < t.size() && t[p] == ',') { p++; continue; }
                if (p < t.size() && t[p] == '}') { p++; return j; }
                throw std::runtime_error("json: expected ',' or '}'");
            }
        }
        if (c == '[')
        {
            Json j = array();
            p++;
            skipWs(t, p);
            if (p < t.size() && t[p] == ']') { p++; return j; }
            while (true)
            {
                j.m_arr.push_back(parseValue(t, p, depth + 1));
                skipWs(t, p);
                if (p < t.size() && t[p] == ',') { p++; continue; }
                if (p < t.size() && t[p] == ']') { p++; return j; }
                throw std::runtime_error("json: expected ',' or ']'");
            }
        }
        if (c == '"') return Json(parseString(t, p));
        if (t.compare(p, 4, "true") == 0) { p += 4; return Json(true); }
        if (t.compare(p, 5, "false") == 0) { p += 5; return Json(false); }
        if (t.compare(p, 4, "null") == 0) { p += 4; return Json(); }
        // number: the JSON grammar ( -? int frac? exp? ) is checked by hand, the digits go through std::from_chars,
        // which ignores the process locale (strtod honours LC_NUMERIC and also takes inf/nan/hex/'+'). Integers
        // without fraction/exponent stay integers (nlohmann keeps them as int64).
        size_t q = p;
        bool integral = true;
        const auto digits = [&]() { const size_t b = q; while (q < t.size() && t[q] >= '0' && t[q] <= '9') q++; return q > b; };
        if (q < t.size() && t[q] == '-') q++;
        if (q < t.size() && t[q] == '0') q++;
        else if (!digits()) throw std::runtime_error("json: invalid value");
        if (q < t.size() && t[q] == '.')
        {
            q++;
            integral = false;
            if (!digits()) throw std::runtime_error("json: digits expected after '.'");
        }
        if (q < t.size() && (t[q] == 'e' || t[q] == 'E'))
        {
            q++;
            integral = false;
            if (q < t.size() && (t[q] == '+' || t[q] == '-')) q++;
            if (!digits()) throw std::runtime_error("json: digits expected in the exponent");
        }
        double d = 0.0;
        const auto res = std::from_chars(t.data() + p, t.data() + q, d);
        if (res.ec == std::errc::result_out_of_range)
            d = t[p] == '-' ? -HUGE_VAL : HUGE_VAL; // like strtod
        else if (res.ec != std::errc() || res.ptr != t.data() + q)
            throw std::runtime_error("json: invalid number");
        p = q;
        Json j(d);
        if (integral) j.m_type = Type::Integer;
        return j;
    }
    static std::string parseString(const std::string& t, size_t& p)
    {
        std::string s;
        p++; // opening quote
        while (p < t.size() && t[p] != '"')
        {
            if (t[p] == '\\' && p + 1 < t.size())
            {
                const char e = t[p + 1];
                p += 2;
                switch (e)
                {
                case 'n': s += '\n'; break;
                case 't': s += '\t'; break;
                case 'r': s += '\r'; break;
                case 'b': s += '\b'; break;
                case 'f': s += '\f'; break;
                case 'u':
                {
                    if (p + 4 > t.size()) throw std::runtime_error("json: bad \\u escape");
                    const unsigned cp = static_cast<unsigned>(std::strtoul(t.substr(p, 4).c_str(), nullptr, 16));
                    p += 4;
                    if (cp < 0x80) s += static_cast<char>(cp);
                    else if (cp < 0x800) { s += static_cast<char>(0xC0 | (cp >> 6)); s += static_cast<char>(0x80 | (cp & 0x3F)); }
                    else { s += static_cast<char>(0xE0 | (cp >> 12)); s += static_cast<char>(0x80 | ((cp >> 6) & 0x3F)); s += static_cast<char>(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: s += e; break; // \" \\ \/
                }
            }
            else
                s += t[p++];
        }
        if (p >= t.size()) throw std::runtime_error("json: unterminated string");
        p++; // closing quote
        return s;
    }
    static void writeNumber(std::string& out, double d)
    {
        if (!std::isfinite(d)) { out += "null"; return; } // nlohmann dumps NaN/inf as null
        if (d == 0.0) { out += std::signbit(d) ? "-0.0" : "0.0"; return; }
        // shortest digit string that round-trips (nlohmann's Grisu2 gives the same digits), laid out
        // like nlohmann: plain decimals for 1e-5 < |d| < 1e15, exponent form outside
        // std::to_chars: shortest round-trip digits, independent of the process locale
        char buf[64];
        auto r = std::to_chars(buf, buf + sizeof(buf), d, std::chars_format::scientific);
        std::string sci(buf, r.ptr);
        const int e10 = std::atoi(sci.c_str() + sci.find('e') + 1);
        if (e10 > -5 && e10 < 15)
        {
            r = std::to_chars(buf, buf + sizeof(buf), d, std::chars_format::fixed);
            std::string fixed(buf, r.ptr);
            if (fixed.find('.') == std::string::npos)
                fixed += ".0";
            out += fixed;
        }
        else
            out += sci;
    }
    static void writeString(std::string& out, const std::string& s)
    {
        out += '"';
        for (const char c : s)
        {
            switch (c)
            {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\t': out += "\\t"; break;
            case '\r': out += "\\r"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            default:
                if (static_cast<unsigned char>(c) < 0x20)
                {
                    char u[8];
                    std::snprintf(u, sizeof(u), "\\u%04x", static_cast<unsigned>(static_cast<unsigned char>(c)));
                    out += u;
                }
                else
                    out += c;
                break;
            }
        }
        out += '"';
    }
    void write(std::string& out, int indent, int depth) const
    {
        const bool pretty = indent >= 0;
        const std::string pad = pretty ? std::string(static_cast<size_t>(indent) * (depth + 1), ' ') : "";
        const std::string padEnd = pretty ? std::string(static_cast<size_t>(indent) * depth, ' ') : "";
        switch (m_type)
        {
        case Type::Null: out += "null"; break;
        case Type::Bool: out += m_bool ? "true" : "false"; break;
        case Type::Integer: out += std::to_string(static_cast<long long>(m_num)); break;
        case Type::Number: writeNumber(out, m_num); break;
        case Type::String: writeString(out, m_str); break;
        case Type::Array:
            if (m_arr.empty()) { out += "[]"; break; }
            out += '[';
            for (size_t i = 0; i < m_arr.size(); i++)
            {
                if (pretty) { out += '\n'; out += pad; }
                m_arr[i].write(out, indent, depth + 1);
                if (i + 1 < m_arr.size()) out += ',';
            }
            if (pretty) { out += '\n'; out += padEnd; }
            out += ']';
            break;
        case Type::Object:
            if (m_obj.empty()) { out += "{}"; break; }
            out += '{';
            {
                size_t i = 0;
                for (const auto& kv : m_obj)
                {
                    if (pretty) { out += '\n'; out += pad; }
                    writeString(out, kv.first);
                    out += pretty ? ": " : ":";
                    kv.second.write(out, indent, depth + 1);
                    if (++i < m_obj.size()) out += ',';
                }
            }
            if (pretty) { out += '\n'; out += padEnd; }
            out += '}';
            break;
        }
    }
};
} // namespace atx
