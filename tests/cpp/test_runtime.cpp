// Scheduler / Sh3Runtime execution-order tests, re-expressed from
// aby3_tests/Sh3RuntimeTests.cpp:15-153 (Task_schedule_test) and :156-267
// (Sh3_Runtime_schedule_test).  Host only: no device is needed or touched.
#include <cstdio>

// -DREF_RUNTIME: the SAME assertions against the reference's own Sh3Runtime / Scheduler, compiled from
// /root/reference with the stand-in headers of oracle/shim (tests/test_ref_parity.py)
#ifdef REF_RUNTIME
#include <aby3/sh3/Sh3Runtime.h>
#else
#include "aby3_b200/sh3/Sh3Runtime.h"
#endif

using namespace aby3;

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static int task_schedule_test() {
    Scheduler rt;
    auto base = rt.nullTask();
    auto task0 = rt.addTask(Type::Round, base);
    auto task1 = rt.addTask(Type::Round, base);
    auto task2 = rt.addTask(Type::Round, std::vector<Task>{task0, task1});
    auto task1b = rt.addTask(Type::Round, base);
    auto task2b = rt.addTask(Type::Round, task2);
    auto close1 = rt.addClosure(task2);
    auto task3 = rt.addTask(Type::Round, close1);

    CHECK(rt.currentTask().mTaskIdx == task0.mTaskIdx); rt.popTask();
    CHECK(rt.currentTask().mTaskIdx == task1.mTaskIdx); rt.popTask();
    CHECK(rt.currentTask().mTaskIdx == task1b.mTaskIdx); rt.popTask();
    CHECK(rt.currentTask().mTaskIdx == task2.mTaskIdx);
    auto task2c = rt.addTask(Type::Round, task2);
    rt.popTask();
    CHECK(rt.currentTask().mTaskIdx == task2b.mTaskIdx); rt.popTask();
    CHECK(rt.currentTask().mTaskIdx == task2c.mTaskIdx); rt.popTask();
    CHECK(rt.currentTask().mTaskIdx == task3.mTaskIdx); rt.popTask();
    CHECK(rt.mTasks.empty());
    return 0;
}

static int runtime_schedule_test() {
    Sh3Runtime rt;
    CommPkg comm;
    rt.init(0, comm);
    int counter = 0;
    bool bad = false;
    auto expect = [&](int v) { if (counter++ != v) bad = true; };
    auto base = rt.noDependencies();

    auto task0 = base.then([&](CommPkg&, Sh3Task self) { expect(0); }, "task0");
    auto task1 = base.then([&](CommPkg&, Sh3Task self) { expect(1); }, "task1");
    auto task2 = (task0 && task1).then([&](CommPkg&, Sh3Task self) {
        expect(2);
        self.then([&](CommPkg&, Sh3Task self) { expect(5); }, "task2-sub1")
            .then([&](CommPkg&, Sh3Task self) { expect(6); }, "task2-sub2");
    }, "task2");
    task2.then([&](Sh3Task self) { expect(4); }, "task2-cont.");
    auto task3 = task2.getClosure().then([&](CommPkg&, Sh3Task self) { expect(7); }, "task3");

    task2.get();
    expect(3);
    task3.get();
    expect(8);

    base.then([&](CommPkg&, Sh3Task self) {
        expect(9);
        self.then([&](CommPkg&, Sh3Task self) { expect(12); });
    });
    base.then([&](CommPkg&, Sh3Task self) {
        expect(10);
        self.then([&](CommPkg&, Sh3Task self) { expect(13); });
    });
    rt.runOneRound();
    expect(11);
    rt.runOneRound();
    expect(14);
    rt.runAll();
    expect(15);
    CHECK(!bad);
    CHECK(counter == 16);
    return 0;
}

static int closure_of_finished_task_is_complete() {
    Sh3Runtime rt;
    CommPkg comm;
    rt.init(0, comm);
    int ran = 0;
    Sh3Task base = rt.noDependencies();
    auto t = base.then([&](CommPkg&, Sh3Task&) { ++ran; });
    t.get();
    CHECK(ran == 1);
    auto c = t.getClosure();
    CHECK(c.isCompleted());
    return 0;
}

static int recursive_get_throws() {
    Sh3Runtime rt;
    CommPkg comm;
    rt.init(0, comm);
    bool threw = false;
    Sh3Task base = rt.noDependencies();
    auto t = base.then([&](CommPkg&, Sh3Task& self) {
        auto inner = self.then([&](CommPkg&, Sh3Task&) {});
        try { inner.get(); } catch (const std::runtime_error&) { threw = true; }
    });
    t.get();
    rt.runAll();
    CHECK(threw);
    return 0;
}

int main() {
    int rc = 0;
    rc |= task_schedule_test();
    rc |= runtime_schedule_test();
    rc |= closure_of_finished_task_is_complete();
    rc |= recursive_get_throws();
    std::printf(rc ? "FAILED\n" : "ALL OK\n");
    return rc;
}
