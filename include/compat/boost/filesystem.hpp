/* boost::filesystem as the sobfu application uses it (src/apps/demo.cpp:178-186,206-221): exists, create_directory, path */
#pragma once
#include <boost/shared_ptr.hpp>
#include <sys/stat.h>
#include <string>
namespace boost {
namespace filesystem {
class path {
public:
    path() {}
    path(const std::string &s) : s_(s) {}
    path(const char *s) : s_(s) {}
    const std::string &string() const { return s_; }
    const char *c_str() const { return s_.c_str(); }
    path operator/(const path &o) const { return path(s_.empty() || s_.back() == '/' ? s_ + o.s_ : s_ + "/" + o.s_); }
private:
    std::string s_;
};
inline bool exists(const path &p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
inline bool is_directory(const path &p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
inline bool create_directory(const path &p) { return ::mkdir(p.c_str(), 0777) == 0; }   /* true when the directory was created */
}  // namespace filesystem
}  // namespace boost
