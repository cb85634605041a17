// TEST INFRASTRUCTURE ONLY: Boost is not in this image.  boost::algorithm::join with the documented behaviour
// (elements separated by `sep`, no leading/trailing separator) for the reference's ReadCounting.cpp.
#pragma once
#include <string>
namespace boost
{
namespace algorithm
{
    template <typename Seq> std::string join(Seq const& parts, std::string const& sep)
    {
        std::string out;
        bool first = true;
        for (auto const& p : parts)
        {
            if (!first)
                out += sep;
            out += p;
            first = false;
        }
        return out;
    }
}
}
