// miplib/cuda/bigvec.hpp - std::vector whose large buffers are recycled between solver objects.
//
// The row store of a bulk-loaded model is hundreds of megabytes (config 3: 384 MB of indices and
// values).  glibc hands such blocks straight to mmap / munmap, so every new Solver pays one page
// fault per 4 KiB when it copies the caller's arrays in - about 0.3 s for config 3, more than the
// host-to-device copy that follows.  Freed blocks of 4 MiB and more are therefore kept in a small
// process-wide cache (at most 4 GiB) and handed to the next vector that asks for a similar size.
#pragma once

#include <cstddef>
#include <cstdlib>
#include <mutex>
#include <new>
#include <type_traits>
#include <utility>
#include <unordered_map>
#include <vector>

namespace miplib {
namespace detail {

struct BlockCache
{
  static constexpr std::size_t kMinBytes = std::size_t(4) << 20;
  static constexpr std::size_t kMaxCached = std::size_t(4) << 30;

  static BlockCache& get()
  {
    static BlockCache* cache = new BlockCache();   // never destroyed: vectors may outlive static teardown
    return *cache;
  }

  void* allocate(std::size_t bytes)
  {
    if (bytes < kMinBytes) return ::operator new(bytes);
    std::size_t const want = (bytes + (std::size_t(2) << 20) - 1) & ~((std::size_t(2) << 20) - 1);
    {
      std::lock_guard<std::mutex> lock(m_mutex);
      std::size_t best = m_free.size();
      for (std::size_t i = 0; i < m_free.size(); ++i)
        if (m_free[i].second >= want && m_free[i].second <= 2 * want &&
            (best == m_free.size() || m_free[i].second < m_free[best].second))
          best = i;
      if (best != m_free.size())
      {
        void* p = m_free[best].first;
        m_cached -= m_free[best].second;
        m_live.emplace(p, m_free[best].second);
        m_free.erase(m_free.begin() + static_cast<std::ptrdiff_t>(best));
        return p;
      }
    }
    void* p = std::aligned_alloc(std::size_t(2) << 20, want);
    if (!p) throw std::bad_alloc();
    std::lock_guard<std::mutex> lock(m_mutex);
    m_live.emplace(p, want);
    return p;
  }

  void deallocate(void* p, std::size_t bytes) noexcept
  {
    if (bytes < kMinBytes) { ::operator delete(p); return; }
    std::lock_guard<std::mutex> lock(m_mutex);
    auto it = m_live.find(p);
    std::size_t const cap = it == m_live.end() ? 0 : it->second;
    if (it != m_live.end()) m_live.erase(it);
    if (cap && m_cached + cap <= kMaxCached && m_free.size() < 32)
    {
      m_free.emplace_back(p, cap);
      m_cached += cap;
    }
    else
      std::free(p);
  }

  private:
  std::mutex m_mutex;
  std::vector<std::pair<void*, std::size_t>> m_free;      // (block, capacity)
  std::unordered_map<void*, std::size_t> m_live;
  std::size_t m_cached = 0;
};

template<class T>
struct RecycleAlloc
{
  using value_type = T;
  RecycleAlloc() = default;
  template<class U> RecycleAlloc(RecycleAlloc<U> const&) {}
  T* allocate(std::size_t n) { return static_cast<T*>(BlockCache::get().allocate(n * sizeof(T))); }
  void deallocate(T* p, std::size_t n) noexcept { BlockCache::get().deallocate(p, n * sizeof(T)); }
  // resize() DEFAULT-initialises (no zero fill for arithmetic types): the bulk ingress grows a vector
  // by hundreds of megabytes and then fills the new tail from several threads, and a value-initialising
  // resize would write every byte once more on one core first
  template<class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value)
  {
    ::new (static_cast<void*>(p)) U;
  }
  template<class U, class... Args> void construct(U* p, Args&&... args)
  {
    ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
  }
  template<class U> bool operator==(RecycleAlloc<U> const&) const { return true; }
  template<class U> bool operator!=(RecycleAlloc<U> const&) const { return false; }
};

}  // namespace detail

template<class T>
using BigVec = std::vector<T, detail::RecycleAlloc<T>>;

}  // namespace miplib
