// Sh3Runtime.h -- the per-party, single-threaded, round-based task scheduler.
// Behaviour follows aby3/Common/Task.h:69-312 and aby3/sh3/Sh3Runtime.{h,cpp}
// (SURVEY 9.1); acceptance = the execution orders pinned by
// aby3_tests/Sh3RuntimeTests.cpp:15-153 and :156-267 (tests/cpp/test_runtime.cpp).
// Nothing here touches the device: a protocol step enqueues kernels and
// event-ordered transfers on the party's stream from inside a task body.
#pragma once
#include <list>
#include <memory>
#include <unordered_map>

#include "Sh3Types.h"

namespace aby3 {

// move-only callable wrapper (the reference uses fu2::unique_function): task
// bodies capture std::future objects by move.
template <typename Sig>
class unique_function;
template <typename R, typename... Args>
class unique_function<R(Args...)> {
    struct Base { virtual ~Base() = default; virtual R call(Args... a) = 0; };
    template <typename F>
    struct Impl : Base {
        F f;
        explicit Impl(F&& x) : f(std::move(x)) {}
        R call(Args... a) override { return f(std::forward<Args>(a)...); }
    };
    std::unique_ptr<Base> mImpl;
public:
    unique_function() = default;
    unique_function(std::nullptr_t) {}
    template <typename F, typename D = typename std::decay<F>::type,
              typename = typename std::enable_if<!std::is_same<D, unique_function>::value &&
                                                 std::is_invocable_r<R, D&, Args...>::value>::type>
    unique_function(F&& f) : mImpl(new Impl<D>(D(std::forward<F>(f)))) {}
    unique_function(unique_function&&) = default;
    unique_function& operator=(unique_function&&) = default;
    explicit operator bool() const { return (bool)mImpl; }
    R operator()(Args... a) { return mImpl->call(std::forward<Args>(a)...); }
};

// ---------------------------------------------------------------- scheduler ----
enum class Type { Round, Continuation };
class Scheduler;
struct Task {
    i64 mTaskIdx = -1;
    Scheduler* mSched = nullptr;
};

class TaskBase {
public:
    Type mType;
    i64 mIdx = -1;
    TaskBase(Type t, u64 idx) : mType(t), mIdx((i64)idx) {}
    std::vector<u64> mUpstream, mDownstream, mClosures;
    static void addUnique(std::vector<u64>& v, u64 x) {
        for (auto y : v) if (y == x) return;
        v.push_back(x);
    }
    void removeUpstream(u64 idx) {
        for (auto& y : mUpstream)
            if (y == idx) { y = mUpstream.back(); mUpstream.pop_back(); return; }
        throw RTE_LOC;
    }
};

class Scheduler {
public:
    i64 mTaskIdx = 0;
    std::unordered_map<u64, TaskBase> mTasks;
    std::list<i64> mReady, mNextRound;

    Task nullTask() { return {-1, this}; }

    // a node that runs (Round) or a pure dependency node; both kinds the runtime
    // creates are registered as Type::Round (Sh3Runtime.cpp:91,118)
    Task addTask(Type t, span<Task> deps) {
        const i64 idx = mTaskIdx++;
        auto it = mTasks.emplace((u64)idx, TaskBase(t, (u64)idx)).first;
        for (auto& d : deps) {
            if (d.mSched != this) throw RTE_LOC;
            if (d.mTaskIdx != -1 && d.mTaskIdx >= idx) throw RTE_LOC;
            auto up = mTasks.find((u64)d.mTaskIdx);
            if (d.mTaskIdx != -1 && up != mTasks.end()) {
                TaskBase::addUnique(up->second.mDownstream, (u64)idx);
                TaskBase::addUnique(it->second.mUpstream, (u64)d.mTaskIdx);
            }
        }
        if (it->second.mUpstream.empty()) queue(mReady, idx);
        return {idx, this};
    }
    Task addTask(Type t, Task dep) { return addTask(t, span<Task>(&dep, 1)); }
    Task addTask(Type t, const std::vector<Task>& deps) {
        return addTask(t, span<Task>(const_cast<Task*>(deps.data()), deps.size()));
    }
    Task addClosure(const std::vector<Task>& deps) {
        return addClosure(span<Task>(const_cast<Task*>(deps.data()), deps.size()));
    }

    // completes when `dep` and everything it transitively spawns have completed;
    // born complete when dep already finished (Task.h:163-200)
    Task addClosure(span<Task> deps) {
        const i64 idx = mTaskIdx++;
        TaskBase x(Type::Continuation, (u64)idx);
        for (auto& d : deps) {
            if (d.mSched != this) throw RTE_LOC;
            if (d.mTaskIdx != -1 && d.mTaskIdx >= idx) throw RTE_LOC;
            auto up = mTasks.find((u64)d.mTaskIdx);
            if (d.mTaskIdx != -1 && up != mTasks.end()) {
                TaskBase::addUnique(up->second.mClosures, (u64)idx);
                TaskBase::addUnique(x.mUpstream, (u64)d.mTaskIdx);
            }
        }
        if (!x.mUpstream.empty()) mTasks.emplace((u64)idx, std::move(x));
        return {idx, this};
    }
    Task addClosure(Task dep) { return addClosure(span<Task>(&dep, 1)); }

    Task currentTask() {
        if (mReady.empty()) std::swap(mReady, mNextRound);
        if (mReady.empty()) throw RTE_LOC;
        return {mReady.front(), this};
    }
    void popTask() {
        if (mReady.empty()) throw RTE_LOC;
        const i64 idx = mReady.front();
        removeTask((u64)idx);
        mReady.pop_front();
    }
    void removeTask(u64 idx) {
        auto task = mTasks.find(idx);
        if (task == mTasks.end()) throw RTE_LOC;
        if (!task->second.mUpstream.empty()) throw RTE_LOC;
        for (auto d : task->second.mDownstream) {
            auto ds = mTasks.find(d);
            if (ds == mTasks.end()) throw RTE_LOC;
            ds->second.removeUpstream(idx);
            if (ds->second.mUpstream.empty()) {
                // children that run go to the NEXT round; dependency-only nodes are ready now
                if (ds->second.mType == Type::Round) queue(mNextRound, (i64)d);
                else queue(mReady, (i64)d);
            }
            // a closure of the finished task now also waits for each of its children
            for (auto c : task->second.mClosures) {
                auto cc = mTasks.find(c);
                if (cc == mTasks.end()) throw RTE_LOC;
                TaskBase::addUnique(ds->second.mClosures, c);
                TaskBase::addUnique(cc->second.mUpstream, d);
            }
        }
        const std::vector<u64> closures = task->second.mClosures;
        for (auto c : closures) {
            auto cc = mTasks.find(c);
            if (cc == mTasks.end()) continue;
            cc->second.removeUpstream(idx);
            if (cc->second.mUpstream.empty()) removeTask(c);
        }
        mTasks.erase(idx);
    }

private:
    void queue(std::list<i64>& l, i64 idx) {
        auto it = mTasks.find((u64)idx);
        if (idx > mTaskIdx || it == mTasks.end() || !it->second.mUpstream.empty()) throw RTE_LOC;
        l.push_back(idx);
    }
};

// ------------------------------------------------------------------ runtime ----
class Sh3Runtime;

class Sh3Task {
public:
    using RoundFunc = unique_function<void(CommPkg& comm, Sh3Task& self)>;
    using ContinuationFunc = unique_function<void(Sh3Task& self)>;
    enum Type { Evaluation, Closure };

    Sh3Task() = default;
    Sh3Task(Sh3Runtime* rt, i64 idx, Type t = Evaluation) : mRuntime(rt), mIdx(idx), mType(t) {}

    Sh3Runtime& getRuntime() const { return *mRuntime; }
    // a task that may run in the round after this one
    Sh3Task then(RoundFunc task);
    Sh3Task then(ContinuationFunc task);
    Sh3Task then(RoundFunc task, std::string name);
    Sh3Task then(ContinuationFunc task, std::string name);
    // fulfilled when this task and everything it spawns are fulfilled
    Sh3Task getClosure();
    std::string& name();
    Sh3Task operator&&(const Sh3Task& o) const;
    Sh3Task operator&=(const Sh3Task& o);
    // run the party's task queue until this task is complete
    void get();
    bool isCompleted();
    bool operator==(const Sh3Task& t) const { return mRuntime == t.mRuntime && mIdx == t.mIdx && mType == t.mType; }
    bool operator!=(const Sh3Task& t) const { return !(*this == t); }

    Sh3Runtime* mRuntime = nullptr;
    i64 mIdx = -1;
    Type mType = Evaluation;
};

inline std::ostream& operator<<(std::ostream& o, const Sh3Task& d) {
    return o << d.mIdx << (d.mType == Sh3Task::Evaluation ? ".E" : ".C");
}

class Sh3TaskBase {
public:
    enum Kind { Empty, And, Round, Continuation };
    Kind mKind = Empty;
    std::string mName;
    Sh3Task::RoundFunc mRound;
    Sh3Task::ContinuationFunc mCont;
};

class Sh3Runtime {
public:
    Sh3Runtime() = default;
    Sh3Runtime(u64 partyIdx, CommPkg& comm) { init(partyIdx, comm); }
    ~Sh3Runtime() {
        if (mSched.mTasks.size())
            std::cout << "~~~~~~~~~~~~~~~~ Runtime not empty!!! ~~~~~~~~~~~~~~~~" << std::endl;
    }

    bool mPrint = false;
    u64 mPartyIdx = (u64)-1;
    CommPkg mComm;
    bool mIsActive = false;

    void init(u64 partyIdx, CommPkg& comm) {
        mPartyIdx = partyIdx;
        mComm = comm;
        mNullTask.mRuntime = this;
        mNullTask.mIdx = -1;
        // the calling thread is this party's thread: bind its device context
        if (comm.mNext.context()) gpu::setCurrent(comm.mNext.context());
    }

    const Sh3Task& noDependencies() const { return mNullTask; }
    operator Sh3Task() const { return noDependencies(); }

    Sh3Task addTask(span<Sh3Task> deps, Sh3Task::RoundFunc&& func, std::string&& name);
    Sh3Task addTask(span<Sh3Task> deps, Sh3Task::ContinuationFunc&& func, std::string&& name);
    Sh3Task addClosure(Sh3Task dep);
    Sh3Task addAnd(span<Sh3Task> deps, std::string&& name);

    void runUntilTaskCompletes(Sh3Task task);
    void runNext();
    void runAll();
    void runOneRound();

    std::unordered_map<u64, Sh3TaskBase> mTasks;
    Scheduler mSched;
    Sh3Task mNullTask;

private:
    std::vector<Task> convert(span<Sh3Task> deps) {
        std::vector<Task> d(deps.size());
        for (u64 i = 0; i < deps.size(); ++i) { d[i].mSched = &mSched; d[i].mTaskIdx = deps[i].mIdx; }
        return d;
    }
};

}  // namespace aby3
