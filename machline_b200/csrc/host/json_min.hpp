// Minimal JSON reader/writer for the MachLine input schema (the reference vendors json-fortran in
// common/json.f90 and looks keys up with dotted paths through common/json_xtnsn.f90:10-15
// `json_xtnsn_get(obj, 'a.b', var, default)`).  Only what the input/report files need.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace mlh {

struct Json {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;  // insertion-ordered

    const Json* find(const std::string& key) const {
        if (type != Object) return nullptr;
        for (auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    // dotted path lookup: "wake_model.append_wake"
    const Json* path(const std::string& p) const {
        const Json* cur = this;
        size_t start = 0;
        while (cur && start <= p.size()) {
            size_t dot = p.find('.', start);
            std::string key = p.substr(start, dot == std::string::npos ? std::string::npos : dot - start);
            cur = cur->find(key);
            if (dot == std::string::npos) break;
            start = dot + 1;
        }
        return cur;
    }
    double get(const std::string& p, double dflt) const {
        const Json* j = path(p);
        return (j && j->type == Number) ? j->num : dflt;
    }
    int get(const std::string& p, int dflt) const {
        const Json* j = path(p);
        return (j && j->type == Number) ? (int)j->num : dflt;
    }
    bool get(const std::string& p, bool dflt) const {
        const Json* j = path(p);
        return (j && j->type == Bool) ? j->b : dflt;
    }
    std::string get(const std::string& p, const std::string& dflt) const {
        const Json* j = path(p);
        return (j && j->type == String) ? j->str : dflt;
    }
    std::string get(const std::string& p, const char* dflt) const { return get(p, std::string(dflt)); }
};

class JsonParser {
public:
    explicit JsonParser(const std::string& text) : s(text), i(0) {}
    Json parse() {
        Json v = value();
        ws();
        return v;
    }

private:
    const std::string& s;
    size_t i;
    void ws() {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
    }
    [[noreturn]] void fail(const char* what) {
        throw std::runtime_error(std::string("JSON parse error: ") + what + " at offset " + std::to_string(i));
    }
    Json value() {
        ws();
        if (i >= s.size()) fail("unexpected end");
        char c = s[i];
        if (c == '{') return object();
        if (c == '[') return array();
        if (c == '"') {
            Json j;
            j.type = Json::String;
            j.str = string();
            return j;
        }
        if (s.compare(i, 4, "true") == 0) {
            i += 4;
            Json j;
            j.type = Json::Bool;
            j.b = true;
            return j;
        }
        if (s.compare(i, 5, "false") == 0) {
            i += 5;
            Json j;
            j.type = Json::Bool;
            j.b = false;
            return j;
        }
        if (s.compare(i, 4, "null") == 0) {
            i += 4;
            return Json();
        }
        return number();
    }
    Json number() {
        const char* start = s.c_str() + i;
        char* end = nullptr;
        double v = std::strtod(start, &end);
        if (end == start) fail("bad number");
        i += (size_t)(end - start);
        Json j;
        j.type = Json::Number;
        j.num = v;
        return j;
    }
    std::string string() {
        std::string out;
        ++i;  // opening quote
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) {
                char e = s[i + 1];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': out += '?'; i += 4; break;
                    default: out += e;
                }
                i += 2;
            } else {
                out += s[i++];
            }
        }
        if (i >= s.size()) fail("unterminated string");
        ++i;
        return out;
    }
    Json array() {
        Json j;
        j.type = Json::Array;
        ++i;
        ws();
        if (i < s.size() && s[i] == ']') {
            ++i;
            return j;
        }
        for (;;) {
            j.arr.push_back(value());
            ws();
            if (i >= s.size()) fail("unterminated array");
            if (s[i] == ',') {
                ++i;
                continue;
            }
            if (s[i] == ']') {
                ++i;
                break;
            }
            fail("expected , or ]");
        }
        return j;
    }
    Json object() {
        Json j;
        j.type = Json::Object;
        ++i;
        ws();
        if (i < s.size() && s[i] == '}') {
            ++i;
            return j;
        }
        for (;;) {
            ws();
            if (i >= s.size() || s[i] != '"') fail("expected key");
            std::string key = string();
            ws();
            if (i >= s.size() || s[i] != ':') fail("expected :");
            ++i;
            j.obj.emplace_back(key, value());
            ws();
            if (i >= s.size()) fail("unterminated object");
            if (s[i] == ',') {
                ++i;
                continue;
            }
            if (s[i] == '}') {
                ++i;
                break;
            }
            fail("expected , or }");
        }
        return j;
    }
};

}  // namespace mlh
