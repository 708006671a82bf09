// Stand-in for Taskflow v3.8.0 (commit d8c49c64, fetched by the reference's
// icicle/backend/cpu/CMakeLists.txt:19-24 and absent offline).
// TEST INFRASTRUCTURE ONLY: lets the reference CPU backend compile into
// oracle/_ref/. It carries no arithmetic; only the three calls the reference
// uses (Taskflow::emplace/clear, Executor::run(...).wait()) are provided.
#pragma once
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

namespace tf {

  class Taskflow
  {
  public:
    template <typename F>
    void emplace(F&& f)
    {
      jobs_.emplace_back(std::forward<F>(f));
    }
    void clear() { jobs_.clear(); }
    std::vector<std::function<void()>>& jobs() { return jobs_; }

  private:
    std::vector<std::function<void()>> jobs_;
  };

  class Executor
  {
  public:
    struct Done {
      void wait() const {}
    };
    explicit Executor(unsigned n = std::thread::hardware_concurrency()) : width_(n ? n : 1) {}

    // Runs every job to completion before returning; wait() is then a no-op.
    Done run(Taskflow& flow)
    {
      auto& jobs = flow.jobs();
      const size_t total = jobs.size();
      if (total == 0) return {};
      std::atomic<size_t> cursor{0};
      auto drain = [&]() {
        for (size_t k = cursor.fetch_add(1); k < total; k = cursor.fetch_add(1))
          jobs[k]();
      };
      const size_t helpers = std::min<size_t>(width_, total) - 1;
      std::vector<std::thread> pool;
      pool.reserve(helpers);
      for (size_t t = 0; t < helpers; ++t)
        pool.emplace_back(drain);
      drain();
      for (auto& th : pool)
        th.join();
      return {};
    }

  private:
    unsigned width_;
  };

} // namespace tf
