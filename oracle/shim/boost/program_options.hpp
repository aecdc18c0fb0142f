// TEST INFRASTRUCTURE ONLY (oracle build).  Non-functional stand-in for
// boost::program_options so that the reference's include/io/configuration.h
// (pulled in by collision.h:8) parses when the reference's L0-L3 headers are
// compiled from /root/reference for oracle/_ref.  io::Config is never
// constructed by the oracle driver, so nothing here does real work.
#pragma once
#include <string>
#include <iostream>
#include <cassert>
#include <cstdint>
#include <cstdlib>

namespace boost { namespace program_options {

template <typename T> struct typed_value {
    typed_value* default_value(const T&) { return this; }
    typed_value* required() { return this; }
};
template <typename T> typed_value<T>* value(T*) { static typed_value<T> v; return &v; }

struct options_description;
struct option_adder {
    option_adder& operator()(const char*, const char*) { return *this; }
    template <typename V> option_adder& operator()(const char*, V*, const char*) { return *this; }
};
struct options_description {
    options_description() {}
    options_description(const char*) {}
    option_adder add_options() { return option_adder(); }
    options_description& add(const options_description&) { return *this; }
};
inline std::ostream& operator<<(std::ostream& os, const options_description&) { return os; }

struct positional_options_description {
    positional_options_description& add(const char*, int) { return *this; }
};

struct variable_value {
    template <typename T> T as() const { return T(); }
};
struct variables_map {
    int count(const char*) const { return 0; }
    variable_value operator[](const char*) const { return variable_value(); }
};

struct parsed_options {};
struct command_line_parser {
    command_line_parser(int, char**) {}
    command_line_parser& options(const options_description&) { return *this; }
    command_line_parser& allow_unregistered() { return *this; }
    command_line_parser& positional(const positional_options_description&) { return *this; }
    parsed_options run() { return parsed_options(); }
};
template <typename C> parsed_options parse_config_file(const C*, const options_description&) {
    return parsed_options();
}
inline void store(const parsed_options&, variables_map&) {}
inline void notify(variables_map&) {}

}} // namespace boost::program_options
