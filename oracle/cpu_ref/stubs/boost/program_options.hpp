// Stand-in for <boost/program_options.hpp>: the one type the reference's host helper header names
// (po::variables_map in checkReportParameters), with the members it touches.  Test infrastructure.
#pragma once
#include <string>
namespace boost { namespace program_options {
struct variable_value {
    bool defaulted() const { return true; }
    bool empty() const { return true; }
    template <class T> T as() const { return T(); }
};
struct variables_map {
    variable_value operator[](const std::string&) const { return variable_value(); }
    size_t count(const std::string&) const { return 0; }
};
} }
