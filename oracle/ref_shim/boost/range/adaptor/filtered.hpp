// TEST INFRASTRUCTURE ONLY: Boost is not in this image.  The one use the reference's PathAligner.cpp makes of
// boost::adaptors::filtered: `range | filtered(pred)` as a lazily filtered view with empty(), front(), begin(), end().
#pragma once
#include <iterator>
#include <utility>
namespace boost
{
namespace adaptors
{
    template <typename Pred> struct filter_holder
    {
        Pred pred;
    };
    template <typename Pred> filter_holder<Pred> filtered(Pred p) { return filter_holder<Pred>{ std::move(p) }; }

    template <typename Range, typename Pred> class filtered_range
    {
        using base_it = decltype(std::begin(std::declval<Range&>()));

    public:
        class iterator
        {
        public:
            using iterator_category = std::forward_iterator_tag;
            using value_type = typename std::iterator_traits<base_it>::value_type;
            using difference_type = std::ptrdiff_t;
            using pointer = typename std::iterator_traits<base_it>::pointer;
            using reference = typename std::iterator_traits<base_it>::reference;
            iterator(base_it it, base_it end, Pred const* pred)
                : it_(it), end_(end), pred_(pred)
            {
                skip();
            }
            reference operator*() const { return *it_; }
            pointer operator->() const { return &*it_; }
            iterator& operator++()
            {
                ++it_;
                skip();
                return *this;
            }
            iterator operator++(int)
            {
                iterator t = *this;
                ++*this;
                return t;
            }
            bool operator==(iterator const& o) const { return it_ == o.it_; }
            bool operator!=(iterator const& o) const { return it_ != o.it_; }

        private:
            void skip()
            {
                while (it_ != end_ && !(*pred_)(*it_))
                    ++it_;
            }
            base_it it_, end_;
            Pred const* pred_;
        };
        filtered_range(Range& r, Pred p)
            : r_(&r), pred_(std::move(p))
        {
        }
        iterator begin() const { return iterator(std::begin(*r_), std::end(*r_), &pred_); }
        iterator end() const { return iterator(std::end(*r_), std::end(*r_), &pred_); }
        bool empty() const { return begin() == end(); }
        auto front() const -> decltype(*std::declval<iterator>()) { return *begin(); }

    private:
        Range* r_;
        Pred pred_;
    };
    template <typename Range, typename Pred> filtered_range<Range, Pred> operator|(Range& r, filter_holder<Pred> h)
    {
        return filtered_range<Range, Pred>(r, std::move(h.pred));
    }
}
}
