// TEST INFRASTRUCTURE ONLY (oracle build).  Non-functional stand-in for
// boost::filesystem, see program_options.hpp in this directory.
#pragma once
#include <string>

namespace boost { namespace filesystem {

struct path {
    std::string s;
    path() {}
    path(const std::string& p) : s(p) {}
};
inline bool exists(const path&) { return true; }
inline bool create_directory(const path&) { return false; }
inline bool remove(const path&) { return false; }

struct directory_entry {
    boost::filesystem::path p;
    const boost::filesystem::path& path() const { return p; }
};
struct directory_iterator {
    directory_entry e;
    directory_iterator() {}
    directory_iterator(const boost::filesystem::path&) {}
    bool operator!=(const directory_iterator&) const { return false; }
    directory_iterator& operator++() { return *this; }
    const directory_entry* operator->() const { return &e; }
};

}} // namespace boost::filesystem
