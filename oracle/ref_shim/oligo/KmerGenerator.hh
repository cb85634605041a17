// TEST INFRASTRUCTURE ONLY (oracle/ref_shim): stand-in for the reference's oligo/KmerGenerator.hh, so that the
// UNMODIFIED src/c++/lib/grm/KmerAligner.cpp compiles here.  The reference's oligo/Kmer.hh is built on Boost.MPL /
// Boost.Preprocessor, which this image does not have; KmerAligner.cpp itself only needs
//   oligo::KmerGenerator<K, unsigned, std::string::const_iterator>  -- successive k-mers of a sequence, 2 bits per base
//                                                                      (A/a 0, C/c 1, G/g 2, T/t 3), k-mers containing
//                                                                      any other character are skipped
//   oligo::bases(kmer)                                               -- only inside a debug operator<<
// (reference: src/c++/include/oligo/KmerGenerator.hh:41-157, Nucleotides.hh:40-46, 59-341).  Everything that decides
// KmerAligner's results apart from the k-mer values -- seeding, the bounded heap of candidates, pickBest, clipping,
// CIGAR building -- is the reference's own code.  This stand-in is pinned by the reference's unit test
// (src/c++/test/test_kmeraligner.cpp:149-191, tests/golden/kmer_aligner.json).
#pragma once
#include <cstddef>
#include <iterator>
#include <string>

namespace oligo
{
static const unsigned BITS_PER_BASE = 2;
static const unsigned int INVALID_OLIGO = 4;

inline unsigned getValue(const char base)
{
    switch (base)
    {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return INVALID_OLIGO;
    }
}

template <typename KmerT> std::string bases(KmerT kmer)
{
    std::string s(sizeof(KmerT) * 8 / BITS_PER_BASE, 'A');
    for (std::size_t i = s.size(); i-- > 0; kmer >>= BITS_PER_BASE)
        s[i] = "ACGT"[kmer & 3];
    return s;
}

template <unsigned kmerLength, class T, typename InputIteratorT, unsigned step = 1> class KmerGenerator
{
public:
    KmerGenerator(const InputIteratorT begin, const InputIteratorT end)
        : next_(begin)
        , end_(end)
        , mask_(kmerLength * BITS_PER_BASE >= sizeof(T) * 8 ? T(~T(0)) : T((T(1) << (kmerLength * BITS_PER_BASE)) - 1))
        , kmer_(0)
        , have_(0)
    {
        static_assert(step == 1, "only step 1 is needed by KmerAligner");
    }

    // the next k-mer without an invalid character and the position of its first base; false at the end of the sequence
    bool next(T& kmer, InputIteratorT& position)
    {
        while (next_ != end_)
        {
            const unsigned v = getValue(*next_);
            ++next_;
            if (v >= INVALID_OLIGO)
            {
                have_ = 0;
                continue;
            }
            kmer_ = T(kmer_ << BITS_PER_BASE) | T(v);
            if (++have_ >= kmerLength)
            {
                kmer = kmer_ & mask_;
                position = next_ - kmerLength;
                return true;
            }
        }
        return false;
    }

private:
    InputIteratorT next_;
    const InputIteratorT end_;
    const T mask_;
    T kmer_;
    unsigned have_;
};
} // namespace oligo
