/*
 * Dependency-free stand-in for the part of boost::program_options that the sobfu application uses to read its .ini
 * files (src/apps/demo.cpp:57-66,84-160,166-171 of the reference): options_description with add_options()(name,
 * value<T>(&dst) | value<T>(), help), parse_config_file(stream, desc), store, notify, variables_map["NAME"].as<T>().
 *
 * Config-file grammar as boost implements it: one `NAME=VALUE` per line, `#` starts a comment, blank lines are skipped,
 * whitespace around the name and the value is trimmed, `[section]` headers prefix the following names with `section.`;
 * an option the description does not know raises unknown_option, a value that does not parse raises invalid_option_value,
 * an option given twice in one file raises multiple_occurrences (options here are non-composing), and across several store()
 * calls the value stored first wins.  Boost is not a dependency of sobfu_b200; with Boost installed, put it first on
 * the include path.
 */
#pragma once
#include <cctype>
#include <istream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

namespace boost {
namespace program_options {

struct error : std::logic_error { explicit error(const std::string &w) : std::logic_error(w) {} };
struct unknown_option : error { explicit unknown_option(const std::string &n) : error("unrecognised option '" + n + "'") {} };
struct invalid_option_value : error {
    invalid_option_value(const std::string &n, const std::string &v) : error("the argument ('" + v + "') for option '" + n + "' is invalid") {}
};
struct multiple_occurrences : error { explicit multiple_occurrences(const std::string &n) : error("option '" + n + "' cannot be specified more than once") {} };
struct invalid_config_file_syntax : error { explicit invalid_config_file_syntax(const std::string &l) : error("the options configuration file contains an invalid line '" + l + "'") {} };

/* holder of one typed value (boost::any in the original) */
struct value_holder {
    virtual ~value_holder() {}
    virtual const std::type_info &type() const = 0;
};
template <typename T>
struct typed_holder : value_holder {
    T v;
    explicit typed_holder(const T &x) : v(x) {}
    const std::type_info &type() const override { return typeid(T); }
};

class value_semantic {
public:
    virtual ~value_semantic() {}
    /* text -> value; false when the text is not a T */
    virtual bool parse(const std::string &text, std::shared_ptr<value_holder> &out) const = 0;
    /* notify(): write the stored value through the pointer given to value<T>(&dst) */
    virtual void notify(const value_holder &v) const = 0;
};

template <typename T>
class typed_value : public value_semantic {
public:
    explicit typed_value(T *store_to) : dst_(store_to) {}
    bool parse(const std::string &text, std::shared_ptr<value_holder> &out) const override {
        T v{};
        if (!convert(text, v)) return false;
        out = std::make_shared<typed_holder<T>>(v);
        return true;
    }
    void notify(const value_holder &h) const override {
        if (dst_) *dst_ = static_cast<const typed_holder<T> &>(h).v;
    }

private:
    template <typename U>
    static bool convert(const std::string &text, U &v) {        // lexical_cast: the whole token has to be consumed
        std::istringstream is(text);
        is >> std::noskipws >> v;
        return !is.fail() && is.peek() == std::char_traits<char>::eof();
    }
    static bool convert(const std::string &text, std::string &v) { v = text; return true; }
    static bool convert(const std::string &text, bool &v) {
        std::string t;
        for (char c : text) t += (char)std::tolower((unsigned char)c);
        if (t == "1" || t == "true" || t == "yes" || t == "on") { v = true; return true; }
        if (t == "0" || t == "false" || t == "no" || t == "off") { v = false; return true; }
        return false;
    }
    T *dst_;
};

template <typename T> typed_value<T> *value() { return new typed_value<T>(nullptr); }
template <typename T> typed_value<T> *value(T *store_to) { return new typed_value<T>(store_to); }

struct option_description {
    std::string name, help;
    std::shared_ptr<const value_semantic> semantic;
};

class options_description;
class options_description_easy_init {
public:
    explicit options_description_easy_init(options_description *owner) : owner_(owner) {}
    options_description_easy_init &operator()(const char *name, const value_semantic *s, const char *help = "");
    options_description_easy_init &operator()(const char *name, const char *help);

private:
    options_description *owner_;
};

class options_description {
public:
    options_description() {}
    explicit options_description(const std::string &caption) : caption_(caption) {}
    options_description_easy_init add_options() { return options_description_easy_init(this); }
    const option_description *find_nothrow(const std::string &name) const {
        for (const auto &o : options_) if (o.name == name) return &o;
        return nullptr;
    }
    const std::vector<option_description> &options() const { return options_; }
    void add(const option_description &o) { options_.push_back(o); }
    const std::string &caption() const { return caption_; }

private:
    std::string caption_;
    std::vector<option_description> options_;
};

inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const value_semantic *s, const char *help) {
    owner_->add(option_description{name, help ? help : "", std::shared_ptr<const value_semantic>(s)});
    return *this;
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const char *help) {
    owner_->add(option_description{name, help ? help : "", nullptr});
    return *this;
}

inline std::ostream &operator<<(std::ostream &os, const options_description &d) {
    if (!d.caption().empty()) os << d.caption() << ":\n";
    for (const auto &o : d.options()) os << "  " << o.name << "  " << o.help << "\n";
    return os;
}

struct basic_option { std::string string_key; std::vector<std::string> value; };
struct parsed_options {
    const options_description *description;
    std::vector<basic_option> options;
};

inline std::string trim_ws(const std::string &s) {
    size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
    return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
}

template <typename charT>
parsed_options parse_config_file(std::basic_istream<charT> &is, const options_description &desc, bool allow_unregistered = false) {
    parsed_options out{&desc, {}};
    std::string line, prefix;
    while (std::getline(is, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        line = trim_ws(line);
        if (line.empty()) continue;
        if (line.front() == '[' && line.back() == ']') {
            prefix = line.substr(1, line.size() - 2);
            if (!prefix.empty() && prefix.back() != '.') prefix += '.';
            continue;
        }
        const size_t eq = line.find('=');
        if (eq == std::string::npos) throw invalid_config_file_syntax(line);
        const std::string name = prefix + trim_ws(line.substr(0, eq)), val = trim_ws(line.substr(eq + 1));
        if (!desc.find_nothrow(name)) {
            if (allow_unregistered) continue;
            throw unknown_option(name);
        }
        out.options.push_back(basic_option{name, {val}});
    }
    return out;
}

class variable_value {
public:
    variable_value() {}
    explicit variable_value(std::shared_ptr<value_holder> v) : v_(std::move(v)) {}
    bool empty() const { return !v_; }
    template <typename T>
    const T &as() const {
        if (!v_ || v_->type() != typeid(T)) throw std::bad_cast();
        return static_cast<const typed_holder<T> &>(*v_).v;
    }
    const value_holder *holder() const { return v_.get(); }

private:
    std::shared_ptr<value_holder> v_;
};

class variables_map : public std::map<std::string, variable_value> {
public:
    const variable_value &operator[](const std::string &name) const {
        static const variable_value none;
        auto it = find(name);
        return it == end() ? none : it->second;
    }
    size_t count(const std::string &name) const { return std::map<std::string, variable_value>::count(name); }
    void notify() {
        for (const auto &kv : sem_) {
            auto it = find(kv.first);
            if (it != end() && kv.second && it->second.holder()) kv.second->notify(*it->second.holder());
        }
    }
    void remember(const std::string &name, std::shared_ptr<const value_semantic> s) { sem_[name] = std::move(s); }

private:
    std::map<std::string, std::shared_ptr<const value_semantic>> sem_;
};

inline void store(const parsed_options &parsed, variables_map &vm) {
    std::map<std::string, int> seen;                              // occurrences inside THIS parse
    for (const auto &o : parsed.options)
        if (++seen[o.string_key] > 1) throw multiple_occurrences(o.string_key);
    for (const auto &o : parsed.options) {
        if (vm.count(o.string_key)) continue;                     // stored by an earlier store(): the first value wins
        const option_description *d = parsed.description->find_nothrow(o.string_key);
        if (!d || !d->semantic) continue;
        std::shared_ptr<value_holder> h;
        if (!d->semantic->parse(o.value.empty() ? std::string() : o.value[0], h)) throw invalid_option_value(o.string_key, o.value.empty() ? "" : o.value[0]);
        vm.insert({o.string_key, variable_value(h)});
        vm.remember(o.string_key, d->semantic);
    }
}
inline void notify(variables_map &vm) { vm.notify(); }

}  // namespace program_options
}  // namespace boost
