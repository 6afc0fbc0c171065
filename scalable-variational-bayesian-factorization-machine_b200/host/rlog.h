// host/rlog.h -- per-iteration TSV log behind `-rlog` (reference src/util/rlog.h:29-91): a header line with the
// registered fields in registration order, then one line per newLine() holding the logged values (default for
// fields not logged since the previous line), default ostream formatting.
#pragma once
#include <algorithm>
#include <map>
#include <ostream>
#include <string>
#include <vector>

namespace svbfm_host {

class RLog {
    std::ostream* out_;
    std::vector<std::string> fields_;
    std::map<std::string, double> dflt_, cur_;

public:
    explicit RLog(std::ostream* out) : out_(out) {}
    void addField(const std::string& name, double dflt) {
        if (std::find(fields_.begin(), fields_.end(), name) != fields_.end()) throw "the field " + name + " already exists";
        fields_.push_back(name);
        dflt_[name] = dflt;
    }
    void log(const std::string& name, double v) { cur_[name] = v; }
    void init() {
        if (out_) {
            for (size_t i = 0; i < fields_.size(); i++) *out_ << fields_[i] << (i + 1 < fields_.size() ? "\t" : "\n");
            out_->flush();
        }
        cur_ = dflt_;
    }
    void newLine() {
        if (!out_) return;
        for (size_t i = 0; i < fields_.size(); i++) *out_ << cur_[fields_[i]] << (i + 1 < fields_.size() ? "\t" : "\n");
        out_->flush();
        cur_ = dflt_;
    }
};

}  // namespace svbfm_host
