#pragma once
// compat: cryptoTools/Common/CLP.h: "-key v1 v2 ..." command lines as a map of string lists, with the accessors the
// reference's mains use (frontend/main.cpp:16-54, aby3-ML/main-linear.cpp:173-180): isSet, get<T>, getOr, getMany,
// getManyOr, set, setDefault, parse.
#include <map>
#include <list>
#include <set>
#include <sstream>
#include "aby3_b200/sh3/Defines.h"
namespace oc {
class CommandLineParserError : public std::exception {};
class CLP {
public:
    CLP() = default;
    CLP(int argc, char** argv) { parse(argc, argv); }
    std::string mProgramName;
    std::map<std::string, std::list<std::string>> mKeyValues;

    void parse(int argc, char const* const* argv) {
        if (argc > 0) mProgramName = argv[0];
        std::string key;
        for (int i = 1; i < argc; ++i) {
            std::string t = argv[i];
            const bool isKey = t.size() > 1 && t[0] == '-' && !(isdigit((unsigned char)t[1]) || t[1] == '.');
            if (isKey) {
                key = t.substr(t.find_first_not_of('-'));
                mKeyValues[key];
            } else {
                if (key.empty()) throw CommandLineParserError();
                mKeyValues[key].push_back(t);
            }
        }
    }
    void set(const std::string& name) { mKeyValues[name]; }
    void setDefault(const std::string& key, const std::string& value) {
        if (!hasValue(key)) mKeyValues[key] = {value};
    }
    void setDefault(const std::vector<std::string>& keys, const std::string& value) {
        if (!hasValue(keys)) setDefault(keys[0], value);
    }
    template <typename T>
    void setDefault(const std::string& key, const T& value) { setDefault(key, toString(value)); }
    template <typename T>
    void setDefault(const std::vector<std::string>& keys, const T& value) { setDefault(keys, toString(value)); }

    bool isSet(const std::string& name) const { return mKeyValues.find(name) != mKeyValues.end(); }
    bool isSet(const std::vector<std::string>& names) const { for (auto& n : names) if (isSet(n)) return true; return false; }
    bool hasValue(const std::string& name) const { auto it = mKeyValues.find(name); return it != mKeyValues.end() && !it->second.empty(); }
    bool hasValue(const std::vector<std::string>& names) const { for (auto& n : names) if (hasValue(n)) return true; return false; }

    template <typename T>
    T get(const std::string& name) const {
        auto it = mKeyValues.find(name);
        if (it == mKeyValues.end() || it->second.empty()) throw CommandLineParserError();
        return fromString<T>(it->second.front());
    }
    template <typename T>
    T get(const std::vector<std::string>& names) const {
        for (auto& n : names) if (hasValue(n)) return get<T>(n);
        throw CommandLineParserError();
    }
    template <typename T>
    T getOr(const std::string& name, const T& alt) const { return hasValue(name) ? get<T>(name) : alt; }
    template <typename T>
    T getOr(const std::vector<std::string>& names, const T& alt) const { return hasValue(names) ? get<T>(names) : alt; }
    template <typename T>
    std::vector<T> getMany(const std::string& name) const {
        std::vector<T> out;
        auto it = mKeyValues.find(name);
        if (it != mKeyValues.end()) for (auto& s : it->second) out.push_back(fromString<T>(s));
        return out;
    }
    template <typename T>
    std::vector<T> getMany(const std::vector<std::string>& names) const {
        for (auto& n : names) if (hasValue(n)) return getMany<T>(n);
        return {};
    }
    template <typename T>
    std::vector<T> getManyOr(const std::string& name, const std::vector<T>& alt) const { return hasValue(name) ? getMany<T>(name) : alt; }
    template <typename T>
    std::vector<T> getManyOr(const std::vector<std::string>& names, const std::vector<T>& alt) const { return hasValue(names) ? getMany<T>(names) : alt; }

private:
    template <typename T>
    static std::string toString(const T& v) { std::ostringstream o; o << v; return o.str(); }
    template <typename T>
    static T fromString(const std::string& s) { std::istringstream i(s); T v{}; i >> v; if (i.fail()) throw CommandLineParserError(); return v; }
};
}  // namespace oc
