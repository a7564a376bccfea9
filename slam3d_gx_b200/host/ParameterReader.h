// ParameterReader.h -- same class, constructor, GetPara() and global pointer as reference
// src/ParameterReader.h:14-81; the yaml-cpp 0.3 dependency (library absent) is replaced by a parser for the
// flat "key: value" files the reference ships (parameters.yaml: `%YAML:1.0` first line, `#` comments,
// scalars only).  GetPara() returns strings and callers atoi/atof them, exactly like the reference.
#pragma once
#include <map>
#include <sstream>
#include <string>
#include "const.h"

class ParameterReader
{
 public:
    ParameterReader(const std::string &para_file);
    // value of a parameter as a string; unknown names print to stderr and return "unknown_para_name"
    // (reference src/ParameterReader.cpp:121-122)
    std::string GetPara(const std::string &para_name);
    bool ok() const { return _ok; }

 protected:
    std::string num2string(double d) { ss.str(""); ss.clear(); ss << d; return ss.str(); }
    std::string num2string(int d) { ss.str(""); ss.clear(); ss << d; return ss.str(); }
    std::string raw(const std::string &key, const std::string &def) const;

    std::stringstream ss;
    std::map<std::string, std::string> _kv;   // every key of the file
    bool _ok;
};

extern ParameterReader *g_pParaReader;
