// Minimal stand-in for the subset of googletest the reference's test/*.cpp use (TEST_F fixtures, ASSERT_/EXPECT_ NEAR, EQ,
// TRUE, LT/GT, RUN_ALL_TESTS).  gtest 1.8.1 is fetched from the network by the reference's build (test/CMakeLists.txt:12-18),
// which is impossible here; callers with a real googletest put it first on the include path instead.
#pragma once
#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {
class Test {
public:
    virtual ~Test() {}
    virtual void SetUp() {}
    virtual void TearDown() {}
    virtual void TestBody() = 0;
    bool failed_ = false;
};
struct Registry {
    struct Entry { std::string name; std::function<Test *()> make; };
    static std::vector<Entry> &all() { static std::vector<Entry> v; return v; }
    static std::string &filter() { static std::string f; return f; }
};
struct Registrar {
    Registrar(const char *suite, const char *name, std::function<Test *()> make) { Registry::all().push_back({std::string(suite) + "." + name, make}); }
};
inline void InitGoogleTest(int *argc, char **argv) {
    for (int i = 1; i < *argc; ++i) {
        std::string a = argv[i];
        if (a.rfind("--gtest_filter=", 0) == 0) Registry::filter() = a.substr(15);
    }
}
inline int RunAllTests() {
    int failed = 0, run = 0;
    for (auto &e : Registry::all()) {
        const std::string &f = Registry::filter();
        if (!f.empty() && f != "*" && e.name.find(f.back() == '*' ? f.substr(0, f.size() - 1) : f) == std::string::npos) continue;
        std::cout << "[ RUN      ] " << e.name << std::endl;
        Test *t = e.make();
        t->SetUp();
        t->TestBody();
        t->TearDown();
        const bool bad = t->failed_;
        delete t;
        std::cout << (bad ? "[  FAILED  ] " : "[       OK ] ") << e.name << std::endl;
        failed += bad ? 1 : 0;
        ++run;
    }
    std::cout << "[==========] " << run << " tests ran, " << failed << " failed." << std::endl;
    return failed ? 1 : 0;
}
}  // namespace testing

#define RUN_ALL_TESTS() ::testing::RunAllTests()
#define TEST_F(fixture, name)                                                                                   \
    class fixture##_##name##_Test : public fixture {                                                            \
        void TestBody() override;                                                                               \
    };                                                                                                          \
    static ::testing::Registrar fixture##_##name##_reg(#fixture, #name, [] { return (::testing::Test *)new fixture##_##name##_Test(); }); \
    void fixture##_##name##_Test::TestBody()
#define TEST(suite, name)                                                                                       \
    class suite##_##name##_Test : public ::testing::Test {                                                      \
        void TestBody() override;                                                                               \
    };                                                                                                          \
    static ::testing::Registrar suite##_##name##_reg(#suite, #name, [] { return (::testing::Test *)new suite##_##name##_Test(); }); \
    void suite##_##name##_Test::TestBody()

#define GTEST_FAIL_(fatal, msg)                                                                                 \
    do {                                                                                                        \
        std::cout << __FILE__ << ":" << __LINE__ << ": Failure\n" << msg << std::endl;                          \
        this->failed_ = true;                                                                                   \
        if (fatal) return;                                                                                      \
    } while (0)
#define GTEST_NEAR_(a, b, tol, fatal)                                                                           \
    do {                                                                                                        \
        const double a__ = (a), b__ = (b), t__ = (tol);                                                         \
        if (!(std::fabs(a__ - b__) <= t__)) {                                                                   \
            std::ostringstream o__;                                                                             \
            o__ << "The difference between " #a " and " #b " is " << std::fabs(a__ - b__) << ", which exceeds " #tol " (" << a__ << " vs " << b__ << ")"; \
            GTEST_FAIL_(fatal, o__.str());                                                                      \
        }                                                                                                       \
    } while (0)
#define GTEST_CMP_(a, b, op, fatal)                                                                             \
    do {                                                                                                        \
        if (!((a)op(b))) {                                                                                      \
            std::ostringstream o__;                                                                             \
            o__ << "Expected: (" #a ") " #op " (" #b "), actual: " << (a) << " vs " << (b);                     \
            GTEST_FAIL_(fatal, o__.str());                                                                      \
        }                                                                                                       \
    } while (0)
#define ASSERT_NEAR(a, b, tol) GTEST_NEAR_(a, b, tol, true)
#define EXPECT_NEAR(a, b, tol) GTEST_NEAR_(a, b, tol, false)
#define ASSERT_EQ(a, b) GTEST_CMP_(a, b, ==, true)
#define EXPECT_EQ(a, b) GTEST_CMP_(a, b, ==, false)
#define ASSERT_NE(a, b) GTEST_CMP_(a, b, !=, true)
#define ASSERT_LT(a, b) GTEST_CMP_(a, b, <, true)
#define ASSERT_LE(a, b) GTEST_CMP_(a, b, <=, true)
#define ASSERT_GT(a, b) GTEST_CMP_(a, b, >, true)
#define ASSERT_GE(a, b) GTEST_CMP_(a, b, >=, true)
#define EXPECT_LT(a, b) GTEST_CMP_(a, b, <, false)
#define EXPECT_GT(a, b) GTEST_CMP_(a, b, >, false)
#define ASSERT_TRUE(c) do { if (!(c)) GTEST_FAIL_(true, "Expected true: " #c); } while (0)
#define EXPECT_TRUE(c) do { if (!(c)) GTEST_FAIL_(false, "Expected true: " #c); } while (0)
#define ASSERT_FALSE(c) do { if (c) GTEST_FAIL_(true, "Expected false: " #c); } while (0)
