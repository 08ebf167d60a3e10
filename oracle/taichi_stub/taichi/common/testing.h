#pragma once
// the core's unit-test macros: the bodies still have to compile, they are never run here
#define TC_STUB_CAT_(a, b) a##b
#define TC_STUB_CAT(a, b) TC_STUB_CAT_(a, b)
#define TC_TEST(name) static void TC_STUB_CAT(tc_stub_test_, __LINE__)()
#define TC_CHECK_EQUAL(a, b, tol) ((void)(a), (void)(b), (void)(tol))
#define TC_CHECK(x) ((void)(x))
