#pragma once
#include <condition_variable>
#include <mutex>
namespace absl {
class Mutex { public: void Lock() { m_.lock(); } void Unlock() { m_.unlock(); } void lock() { m_.lock(); } void unlock() { m_.unlock(); }
  void AssertHeld() const {} std::mutex m_; };
class MutexLock { public: explicit MutexLock(Mutex* m) : m_(m) { m_->Lock(); } explicit MutexLock(Mutex& m) : m_(&m) { m_->Lock(); } ~MutexLock() { m_->Unlock(); } private: Mutex* m_; };
class CondVar { public: void Wait(Mutex* m) { std::unique_lock<std::mutex> l(m->m_, std::adopt_lock); cv_.wait(l); l.release(); }
  void Signal() { cv_.notify_one(); } void SignalAll() { cv_.notify_all(); } private: std::condition_variable cv_; };
}
