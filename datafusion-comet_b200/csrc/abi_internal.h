// abi_internal.h -- handle layouts and the exception -> error-code boundary shared by the translation units that implement
// include/comet_b200.h (abi.cpp, exchange.cpp).
#pragma once
#include "../../include/comet_b200.h"

#include "exec.h"
#include "jit.h"
#include "plan.h"

#include <cstdio>
#include <string>

struct cb200_table {
    std::shared_ptr<cb200::DeviceTable> t;
};

struct cb200_plan {
    cb200::OperatorP op;
    cb200::ExecContext ctx;
    cb200::PlanInputs inputs;
    cb200::ExecNodeP root;
    cb200::Batch last; // keeps device results alive for cb200_execute_device
    bool started = false, finished = false;
    int64_t export_pos = 0;      // cb200_execute hands `last` out in slices of at most spark.comet.batchSize rows: next row to export
    bool export_pending = false; // ... and whether rows of `last` are still waiting
    int partition = 0, partition_count = 1;
};


inline cb200::ExecContext& cb200_plan_ctx(cb200_plan* p) { return p->ctx; }
inline cb200::Batch& cb200_plan_last(cb200_plan* p) { return p->last; }
inline cb200_table* cb200_table_wrap(std::shared_ptr<cb200::DeviceTable> t) {
    auto* h = new cb200_table();
    h->t = std::move(t);
    return h;
}

inline void set_error(cb200_error* e, int code, const std::string& cls, const std::string& msg) {
    if (!e) return;
    e->code = code;
    snprintf(e->error_class, sizeof(e->error_class), "%s", cls.c_str());
    snprintf(e->message, sizeof(e->message), "%s", msg.c_str());
}
inline void clear_error(cb200_error* e) {
    if (e) { e->code = 0; e->error_class[0] = 0; e->message[0] = 0; }
}

template <typename F> auto cb200_guarded(cb200_error* err, F&& f, decltype(f()) on_error) -> decltype(f()) {
    clear_error(err);
    try {
        return f();
    } catch (const cb200::Unsupported& e) {
        set_error(err, CB200_ERR_UNSUPPORTED, "", e.what());
    } catch (const cb200::PlanError& e) {
        set_error(err, CB200_ERR_PLAN, "", e.what());
    } catch (const cb200::JitError& e) {
        set_error(err, CB200_ERR_JIT, "", e.what());
    } catch (const cb200::ExecError& e) {
        set_error(err, e.code >= 10 ? CB200_ERR_SPARK : e.code, e.error_class, e.what());
    } catch (const std::exception& e) { // the reference turns panics into a pending exception (errors.rs:832-850)
        set_error(err, CB200_ERR_PLAN, "", std::string("native panic: ") + e.what());
    } catch (...) {
        set_error(err, CB200_ERR_PLAN, "", "native panic: unknown exception");
    }
    return on_error;
}

