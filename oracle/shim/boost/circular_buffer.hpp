/* Shim for <boost/circular_buffer.hpp> (Boost is not installed / not part of the reference tree).
 * Only the members messageQueue.h uses: ctor(capacity), full, empty, push_front, back, pop_back,
 * rbegin, rend, reverse_iterator.  A fixed-capacity deque has the same observable behaviour for
 * those calls (the reference never pushes into a full buffer: messageQueue.h:82-86,266-271). */
#ifndef SCN_SHIM_BOOST_CIRCULAR_BUFFER_HPP_
#define SCN_SHIM_BOOST_CIRCULAR_BUFFER_HPP_
// the real header pulls these in transitively; messageQueue.h relies on that
#include <algorithm>
#include <cassert>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <string>
namespace boost {
template <typename T>
class circular_buffer {
  std::deque<T> d_;
  std::size_t cap_;
 public:
  typedef typename std::deque<T>::reverse_iterator reverse_iterator;
  explicit circular_buffer(std::size_t capacity) : cap_(capacity) {}
  bool full() const { return d_.size() >= cap_; }
  bool empty() const { return d_.empty(); }
  std::size_t size() const { return d_.size(); }
  void push_front(const T& v) { if (full() && !d_.empty()) d_.pop_back(); d_.push_front(v); }
  T& back() { return d_.back(); }
  void pop_back() { d_.pop_back(); }
  reverse_iterator rbegin() { return d_.rbegin(); }
  reverse_iterator rend() { return d_.rend(); }
};
}
#endif
