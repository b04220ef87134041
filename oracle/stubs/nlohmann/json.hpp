// TEST INFRASTRUCTURE (oracle/): the sliver of nlohmann::json that the reference's Int8Quan(conf, num_source)
// constructor uses (int8_quan.cc:28-39): stream extraction, size(), operator[](string) and conversion to
// std::string.  Parses objects whose values are strings or objects -- the shape of the reference's
// model-conf files ({"0": {"model_path": "..."}, ...}).  nlohmann-json is an un-vendored dependency.
#pragma once
#include <istream>
#include <iterator>
#include <map>
#include <stdexcept>
#include <string>

namespace nlohmann {

class json {
public:
    size_t size() const { return is_obj_ ? obj_.size() : 1; }
    const json& operator[](const std::string& key) const {
        std::map<std::string, json>::const_iterator it = obj_.find(key);
        if (it == obj_.end()) throw std::out_of_range("json stub: key not found: " + key);
        return it->second;
    }
    operator std::string() const { return str_; }
    friend std::istream& operator>>(std::istream& in, json& j) {
        std::string t((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        size_t p = 0;
        j = parse(t, p);
        return in;
    }

private:
    static void ws(const std::string& t, size_t& p) { while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\t' || t[p] == '\r')) p++; }
    static std::string str(const std::string& t, size_t& p) {
        if (t[p] != '"') throw std::runtime_error("json stub: string expected");
        std::string s;
        for (p++; p < t.size() && t[p] != '"'; p++) {
            if (t[p] == '\\' && p + 1 < t.size()) p++;
            s += t[p];
        }
        p++;
        return s;
    }
    static json parse(const std::string& t, size_t& p) {
        json j;
        ws(t, p);
        if (p < t.size() && t[p] == '{') {
            j.is_obj_ = true;
            p++;
            for (;;) {
                ws(t, p);
                if (p >= t.size()) throw std::runtime_error("json stub: unterminated object");
                if (t[p] == '}') { p++; break; }
                if (t[p] == ',') { p++; continue; }
                const std::string k = str(t, p);
                ws(t, p);
                if (t[p] != ':') throw std::runtime_error("json stub: ':' expected");
                p++;
                j.obj_[k] = parse(t, p);
            }
        } else {
            j.str_ = str(t, p);
        }
        return j;
    }
    bool is_obj_ = false;
    std::map<std::string, json> obj_;
    std::string str_;
};

}  // namespace nlohmann
