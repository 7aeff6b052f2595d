// NCCL, bound at run time.  The engine uses NCCL only when a problem spans several GPUs of one
// process (rssync_create_multi): the library is dlopen'ed the first time such a problem is created
// (libnccl.so.2 -- in a PyTorch process that is the copy torch has already loaded), so single-GPU
// users need no NCCL at all and there is no link-time dependency.  Only the handful of entry
// points used are declared; the types are the ABI's (nccl.h): communicators are opaque pointers,
// results and data types are ints.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>

namespace rs {

struct Nccl {
    using Comm = void*;
    static constexpr int kChar = 0;  // ncclInt8 / ncclChar
    int (*CommInitAll)(Comm*, int, const int*) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* handle = nullptr;

    // nullptr when the library or one of the symbols is missing
    static const Nccl* get() {
        static const Nccl inst = load();
        return inst.handle ? &inst : nullptr;
    }

   private:
    static Nccl load() {
        Nccl n;
        void* h = nullptr;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) return n;
        bool ok = true;
        auto sym = [&](const char* s) {
            void* p = dlsym(h, s);
            ok = ok && p != nullptr;
            return p;
        };
        n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
        n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
        n.Broadcast = reinterpret_cast<decltype(n.Broadcast)>(sym("ncclBroadcast"));
        n.AllGather = reinterpret_cast<decltype(n.AllGather)>(sym("ncclAllGather"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
        if (ok) n.handle = h;
        return n;
    }
};

}  // namespace rs
