// TEST INFRASTRUCTURE ONLY (oracle/). Minimal stand-in for the reference's
// common/Error.hh so that src/c++/lib/grm/GraphAligner.cpp compiles without
// spdlog/Boost (the real header pulls boost/filesystem.hpp, Error.hh:169).
// Only what GraphAligner.cpp uses: LOG()->trace(...) and assert().
#pragma once
#include <cassert>
#include <stdexcept>
#include <string>

namespace pgref_shim
{
struct NullLogger
{
    template <typename... A> void trace(A const&...) const {}
    template <typename... A> void debug(A const&...) const {}
    template <typename... A> void info(A const&...) const {}
    template <typename... A> void warn(A const&...) const {}
    template <typename... A> void error(A const&...) const {}
    template <typename... A> void critical(A const&...) const {}
};
}

static inline pgref_shim::NullLogger* LOG()
{
    static pgref_shim::NullLogger l;
    return &l;
}
