// sm/Runtime.h -- runtime controls of the device path.  NOT part of the reference's surface: the
// reference is synchronous CPU code with nothing to control.  Everything here is opt-in; a program
// that never includes or calls it behaves exactly like the reference (every operator complete on
// return, one GPU).
//
//   sm::async_scope   RAII: inside the scope SMArray operators ENQUEUE their kernels and return
//                     ("async result hand-off"): a loop like the reference's
//                     own benchmark/add.cpp:21-29 (`auto c = a + b;` 10^6 elements, repeated) then costs
//                     a kernel launch per iteration instead of a launch plus a host-device round
//                     trip.  Results are complete when the scope ends, after sm::sync(), and -- without
//                     any call -- before the host reads an array through operator(), toString or
//                     operator% (the headers wait in those places; raw reads through `.data` are the
//                     caller's business: call sm::sync() first).
//   sm::set_devices   spread every operator on large arrays over several GPUs of the box by flat
//                     output range; the program still only writes `a + b`.
#pragma once
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <vector>

#include <smb200.h>

namespace sm {
    inline void sync() {
        if (smb_sync() != SMB_OK) throw std::runtime_error(std::string("smb200: ") + smb_last_error());
    }

    class async_scope {
    public:
        async_scope() : was_(smb_get_option(SMB_OPT_ASYNC)) { smb_set_option(SMB_OPT_ASYNC, 1); }
        async_scope(const async_scope &) = delete;
        async_scope &operator=(const async_scope &) = delete;
        // leaving the outermost scope waits for everything enqueued inside it
        ~async_scope() { if (!was_) smb_set_option(SMB_OPT_ASYNC, 0); }
    private:
        int64_t was_;
    };

    inline void set_devices(const std::vector<int> &devices) {
        if (smb_set_devices(devices.data(), static_cast<int>(devices.size())) != SMB_OK)
            throw std::runtime_error(std::string("smb200: ") + smb_last_error());
    }
    inline void set_devices(std::initializer_list<int> devices) { set_devices(std::vector<int>(devices)); }
    // every GPU of the box
    inline void use_all_devices() {
        std::vector<int> all(static_cast<size_t>(smb_device_count()));
        for (size_t i = 0; i < all.size(); ++i) all[i] = static_cast<int>(i);
        set_devices(all);
    }
    inline std::vector<int> devices() {
        int d[64];
        const int n = smb_get_devices(d, 64);
        return std::vector<int>(d, d + (n < 64 ? n : 64));
    }
} // namespace sm
