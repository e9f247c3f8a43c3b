// Sh3Runtime.cpp -- see Sh3Runtime.h.  Semantics of aby3/sh3/Sh3Runtime.cpp:13-303.
#include "Sh3Runtime.h"

namespace aby3 {

gpu::Context*& gpu::currentSlot() {
    static thread_local gpu::Context* c = nullptr;
    return c;
}

}  // namespace aby3

#include <random>
namespace oc {
block sysRandomSeed() {
    std::random_device rd;
    const u64 a = ((u64)rd() << 32) | rd(), b = ((u64)rd() << 32) | rd();
    return toBlock(a, b);
}
}  // namespace oc

namespace aby3 {

Sh3Task Sh3Task::then(RoundFunc task) { return getRuntime().addTask({this, 1}, std::move(task), {}); }
Sh3Task Sh3Task::then(ContinuationFunc task) { return getRuntime().addTask({this, 1}, std::move(task), {}); }
Sh3Task Sh3Task::then(RoundFunc task, std::string name) { return getRuntime().addTask({this, 1}, std::move(task), std::move(name)); }
Sh3Task Sh3Task::then(ContinuationFunc task, std::string name) { return getRuntime().addTask({this, 1}, std::move(task), std::move(name)); }

Sh3Task Sh3Task::getClosure() { return getRuntime().addClosure(*this); }

std::string& Sh3Task::name() {
    auto it = mRuntime->mTasks.find((u64)mIdx);
    if (it == mRuntime->mTasks.end()) throw RTE_LOC;
    return it->second.mName;
}

Sh3Task Sh3Task::operator&&(const Sh3Task& o) const {
    std::array<Sh3Task, 2> deps{*this, o};
    return getRuntime().addAnd({deps.data(), 2}, {});
}
Sh3Task Sh3Task::operator&=(const Sh3Task& o) {
    *this = *this && o;
    return *this;
}

void Sh3Task::get() { getRuntime().runUntilTaskCompletes(*this); }

bool Sh3Task::isCompleted() { return mRuntime->mSched.mTasks.find((u64)mIdx) == mRuntime->mSched.mTasks.end(); }

Sh3Task Sh3Runtime::addTask(span<Sh3Task> deps, Sh3Task::RoundFunc&& func, std::string&& name) {
    if (!func) { std::cout << "empty task (round function)" << std::endl; throw RTE_LOC; }
    auto d = convert(deps);
    auto tt = mSched.addTask(aby3::Type::Round, span<Task>(d.data(), d.size()));
    auto& t = mTasks[(u64)tt.mTaskIdx];
    t.mKind = Sh3TaskBase::Round;
    t.mRound = std::move(func);
    t.mName = std::move(name);
    return {this, tt.mTaskIdx};
}

Sh3Task Sh3Runtime::addTask(span<Sh3Task> deps, Sh3Task::ContinuationFunc&& func, std::string&& name) {
    if (!func) { std::cout << "empty task (round function)" << std::endl; throw RTE_LOC; }
    auto d = convert(deps);
    // NB: continuation bodies are scheduled like round bodies (Sh3Runtime.cpp:118)
    auto tt = mSched.addTask(aby3::Type::Round, span<Task>(d.data(), d.size()));
    auto& t = mTasks[(u64)tt.mTaskIdx];
    t.mKind = Sh3TaskBase::Continuation;
    t.mCont = std::move(func);
    t.mName = std::move(name);
    return {this, tt.mTaskIdx};
}

Sh3Task Sh3Runtime::addClosure(Sh3Task dep) {
    Task dd;
    dd.mSched = &mSched;
    dd.mTaskIdx = dep.mIdx;
    auto tt = mSched.addClosure(dd);
    return {this, tt.mTaskIdx};
}

Sh3Task Sh3Runtime::addAnd(span<Sh3Task> deps, std::string&& name) {
    auto d = convert(deps);
    auto tt = mSched.addTask(aby3::Type::Round, span<Task>(d.data(), d.size()));
    auto& t = mTasks[(u64)tt.mTaskIdx];
    t.mKind = Sh3TaskBase::And;
    t.mName = std::move(name);
    return {this, tt.mTaskIdx};
}

void Sh3Runtime::runUntilTaskCompletes(Sh3Task task) {
    while (!task.isCompleted()) runNext();
}

void Sh3Runtime::runAll() {
    while (mTasks.size()) runNext();
}

void Sh3Runtime::runOneRound() {
    if (mSched.mTasks.empty()) return;
    mSched.currentTask();
    while (mSched.mReady.size()) runNext();
}

void Sh3Runtime::runNext() {
    if (mIsActive)
        throw std::runtime_error("The runtime is currently running a different task. Do not call Sh3Task.get() recursively. " LOCATION);
    auto tt = mSched.currentTask();
    auto it = mTasks.find((u64)tt.mTaskIdx);
    if (it == mTasks.end()) throw RTE_LOC;
    Sh3Task self{this, tt.mTaskIdx, Sh3Task::Evaluation};
    mIsActive = true;
    try {
        if (it->second.mKind == Sh3TaskBase::Round) it->second.mRound(mComm, self);
        else if (it->second.mKind == Sh3TaskBase::Continuation) it->second.mCont(self);
    } catch (...) {
        mIsActive = false;
        throw;
    }
    mIsActive = false;
    // NCCL transport: everything the task queued (sends, posted receives) goes out as one group
    mComm.mNext.flush();
    mComm.mPrev.flush();
    mTasks.erase((u64)tt.mTaskIdx);
    mSched.popTask();
}

}  // namespace aby3
