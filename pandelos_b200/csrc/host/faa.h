// Line access to a .faa file for the native hosts (pangenes, calculate_k): the file is mapped, lines are handed out as
// [begin, end) byte ranges without copies.  Line terminators are "\n", "\r\n" and a lone "\r", which is what both
// readers of the reference accept (Java BufferedReader.readLine in PangeneIData.java:38, Python's universal newlines
// in calculate_k.py:8); a last line without terminator counts.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace pd_host {

class MappedFile {
public:
    explicit MappedFile(const std::string& path) {
        fd_ = open(path.c_str(), O_RDONLY);
        if (fd_ < 0) return;
        struct stat st;
        if (fstat(fd_, &st) != 0) return;
        n_ = static_cast<size_t>(st.st_size);
        if (n_ == 0) {
            ok_ = true;
            return;
        }
        void* m = mmap(nullptr, n_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m != MAP_FAILED) {
            p_ = static_cast<const char*>(m);
            mapped_ = true;
            madvise(m, n_, MADV_SEQUENTIAL);
            ok_ = true;
            return;
        }
        // not mappable (a pipe, some network file systems): read it
        buf_.resize(n_);
        size_t got = 0;
        while (got < n_) {
            const ssize_t r = read(fd_, &buf_[got], n_ - got);
            if (r <= 0) return;
            got += static_cast<size_t>(r);
        }
        p_ = buf_.data();
        ok_ = true;
    }
    ~MappedFile() {
        if (mapped_) munmap(const_cast<char*>(p_), n_);
        if (fd_ >= 0) close(fd_);
    }
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    bool ok() const { return ok_; }
    const char* data() const { return p_; }
    size_t size() const { return n_; }

private:
    int fd_ = -1;
    const char* p_ = nullptr;
    size_t n_ = 0;
    bool mapped_ = false, ok_ = false;
    std::vector<char> buf_;
};

// fn(index, begin, end) for every line, terminator excluded
template <class F>
void for_each_line(const char* p, size_t n, F fn) {
    const char* const end = p + n;
    size_t index = 0;
    while (p < end) {
        const char* nl = static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(end - p)));
        const char* stop = nl ? nl : end;
        // a lone '\r' inside [p, stop) also ends a line; "\r\n" is one terminator
        const char* q = p;
        for (;;) {
            const char* cr = static_cast<const char*>(memchr(q, '\r', static_cast<size_t>(stop - q)));
            if (!cr) break;
            if (cr + 1 == stop && nl) break;  // the '\r' of "\r\n": handled below
            fn(index++, q, cr);
            q = cr + 1;
        }
        const char* e = stop;
        if (nl && e > q && e[-1] == '\r') e--;
        if (nl || q < end) fn(index++, q, e);  // nothing follows a terminator that ends the file
        p = nl ? nl + 1 : end;
    }
}

}  // namespace pd_host
