// gs_pack.hpp -- host-side 2-bit packer for read batches (the bytes that cross PCIe).
//
// gs_match_submit / gs_filter_submit take the reads as the reference's parser leaves them: ASCII, one byte per base
// (C/fastq/AbstractFastqReader.java:288-368).  Copied as they are, a batch costs 1 byte per base on the host->device link, which
// is what bounds the end-to-end rate (DESIGN.md "PCIe").  The packer turns the batch into exactly the two streams the label
// kernel stages into shared memory anyway:
//   codes[w]  u64: bases 32w .. 32w+31 as 2-bit codes C=0 G=1 A=2 T=3 (C/util/CGAT.java:66-69), first base in the top two bits
//   valid[w]  u32: bit i = base 32w+i is one of the upper-case letters CGAT (everything else breaks the k-mer window, :60-69)
// = 0.375 bytes per base.  Positions past the last base are invalid.  The result is a pure function of the input bytes, so the
// kernels produce the same labels from either form (tests/test_gpu_match.py "packed submit").
#pragma once
#include <cstddef>
#include <cstdint>

namespace gsp {

// Packs n bases (n need not be a multiple of 32; the tail of the last word is invalid) -- single thread, whole words
// [0, ceil(n / 32)).  Picks the AVX-512 or AVX2 body at run time when the CPU has it.
void pack_range(const uint8_t* bases, uint64_t n, uint64_t* codes, uint32_t* valid);

// A small persistent pool: pack() splits the batch into word-aligned pieces that the workers claim from a shared counter;
// the calling thread works too and returns when the whole batch is packed.
class Packer {
public:
    explicit Packer(int threads);   // threads <= 0: one per CPU this process may run on, at most 32
    ~Packer();
    int threads() const;
    void pack(const uint8_t* bases, uint64_t n, uint64_t* codes, uint32_t* valid);
    Packer(const Packer&) = delete;
    Packer& operator=(const Packer&) = delete;
private:
    struct Impl;
    Impl* p;
};

const char* pack_isa();  // "avx512", "avx2" or "scalar": which body pack_range uses on this CPU

}  // namespace gsp
