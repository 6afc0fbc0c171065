// host/cmdline.h -- `-name value` command line, same surface as the reference's CMDLine
// (reference src/util/cmdline.h:29-197): one or two leading dashes, a flag followed by another flag (or by
// nothing) has the empty value, a repeated flag or an unregistered flag is an error reported as a thrown
// std::string, list values are split on ';' and ','.
#pragma once
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>
#include <vector>

namespace svbfm_host {

class CmdLine {
    std::map<std::string, std::string> help_, value_;

    static bool strip_dashes(std::string& s) {
        if (s.empty() || s[0] != '-') return false;
        s.erase(0, (s.size() > 1 && s[1] == '-') ? 2 : 1);
        return true;
    }

public:
    CmdLine(int argc, char** argv) {
        for (int i = 1; i < argc; i++) {
            std::string name(argv[i]);
            if (!strip_dashes(name)) throw "cannot parse " + name;
            if (value_.count(name)) throw "the parameter " + name + " is already specified";
            std::string val;
            if (i + 1 < argc) {
                std::string next(argv[i + 1]);
                std::string probe = next;
                if (!strip_dashes(probe)) { val = next; i++; }
            }
            value_[name] = val;
        }
    }
    const std::string& reg(const std::string& name, const std::string& help) { help_[name] = help; return help_.find(name)->first; }
    void check() const {
        for (auto& kv : value_)
            if (!help_.count(kv.first)) throw "the parameter " + kv.first + " does not exist";
    }
    bool has(const std::string& name) const { return value_.count(name) != 0; }
    void set(const std::string& name, const std::string& v) { value_[name] = v; }
    std::string get(const std::string& name, const std::string& dflt = "") const {
        auto it = value_.find(name);
        return it == value_.end() ? dflt : it->second;
    }
    double get_double(const std::string& name, double dflt) const { return has(name) ? atof(get(name).c_str()) : dflt; }
    long get_int(const std::string& name, long dflt) const { return has(name) ? atol(get(name).c_str()) : dflt; }
    std::vector<std::string> get_list(const std::string& name) const {
        std::vector<std::string> out;
        std::string s = get(name), cur;
        for (char c : s) {
            if (c == ';' || c == ',') { if (!cur.empty()) out.push_back(cur); cur.clear(); }
            else cur.push_back(c);
        }
        if (!cur.empty()) out.push_back(cur);
        return out;
    }
    std::vector<double> get_doubles(const std::string& name) const {
        std::vector<double> out;
        for (auto& s : get_list(name)) out.push_back(atof(s.c_str()));
        return out;
    }
    std::vector<int> get_ints(const std::string& name) const {
        std::vector<int> out;
        for (auto& s : get_list(name)) out.push_back(atoi(s.c_str()));
        return out;
    }
    void print_help() const {
        for (auto& kv : help_) {
            std::string line = "-" + kv.first;
            while (line.size() < 16) line.push_back(' ');
            std::cout << line << kv.second << std::endl;
        }
    }
};

}  // namespace svbfm_host
