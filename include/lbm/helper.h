// include/lbm/helper.h -- the two utilities of the reference's include/helper.h:7-15.
#pragma once
#include <fstream>
#include <memory>
#include <string>
#include <utility>

template <typename T, typename... Args>
std::unique_ptr<T> make_unique(Args&&... args)
{
    return std::unique_ptr<T>(new T(std::forward<Args>(args)...));
}

inline bool file_exists(const std::string& name)
{
    std::ifstream probe(name.c_str());
    return probe.good();
}
