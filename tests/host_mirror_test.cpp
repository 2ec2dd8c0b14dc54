// Exercises the C++ host mirror (rust-lz-fear_b200/host/lz_fear.hpp) the way the reference's own
// unit tests use the crate (src/lib.rs:24-106, src/raw/decompress.rs:153-175, tests/issue-15.rs shape).
#include <cstdio>
#include <sstream>
#include <string>
#include <vector>

#include "../rust-lz-fear_b200/host/lz_fear.hpp"

using namespace lz_fear;

static int fails = 0;
#define CHECK(x) do { if (!(x)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); fails++; } } while (0)

static std::vector<uint8_t> compress(const std::string& s) {              // src/lib.rs:24-32
    std::ostringstream w;
    if (s.size() <= 0xFFFF) { raw::U16Table t; raw::compress2((const uint8_t*)s.data(), s.size(), 0, t, w); }
    else { raw::U32Table t; raw::compress2((const uint8_t*)s.data(), s.size(), 0, t, w); }
    const std::string o = w.str();
    return std::vector<uint8_t>(o.begin(), o.end());
}
static void inverse(const std::string& s) {                               // src/lib.rs:35-41
    const std::vector<uint8_t> c = compress(s);
    std::vector<uint8_t> out;
    raw::decompress_raw(c.data(), c.size(), nullptr, 0, out, size_t(1) << 31);
    CHECK(std::string(out.begin(), out.end()) == s);
}

int main() {
    inverse("to live or not to live");
    inverse("There is nothing either good or bad, but thinking makes it so.");
    inverse("as6yhol.;jrew5tyuikbfewedfyjltre22459ba");
    inverse("ahhd"); inverse("x"); inverse(""); inverse(std::string(13, '\0'));
    const std::string s = "The Read trait allows for reading bytes from a source. Implementors of the Read trait are called "
                          "'readers'. Readers are defined by one required method, read().";
    inverse(s);
    CHECK(compress(s).size() < s.size());
    {   // decode KATs, src/raw/decompress.rs:153-175
        const uint8_t k1[] = {0x11, 'a', 1, 0, 0x22, 'b', 'c', 2, 0};
        std::vector<uint8_t> out;
        raw::decompress_raw(k1, sizeof(k1), nullptr, 0, out, 1 << 20);
        CHECK(std::string(out.begin(), out.end()) == "aaaaaabcbcbcbc");
        const uint8_t k2[] = {0x10, 'a', 2, 0};
        bool threw = false;
        try { std::vector<uint8_t> o2; raw::decompress_raw(k2, sizeof(k2), nullptr, 0, o2, 1 << 20); }
        catch (const raw::DecodeError& e) { threw = e.kind == raw::DecodeError::InvalidDeduplicationOffset; }
        CHECK(threw);
        uint8_t small[4];
        threw = false;
        try { raw::compress_into((const uint8_t*)s.data(), s.size(), small, sizeof(small)); } catch (const raw::WriterFull&) { threw = true; }
        CHECK(threw);
    }
    {   // compress2 with history and a table that lives across calls (src/raw/compress/mod.rs:165-170): the second half
        // repeats the first (pseudo-random bytes), so it compresses only through matches that reach into the history
        std::string half;
        uint32_t x = 12345;
        for (int i = 0; i < 40000; i++) { x = x * 1664525u + 1013904223u; half += (char)(x >> 24); }
        const std::string both = half + half;
        raw::U32Table t;
        std::ostringstream w1, w2;
        raw::compress2((const uint8_t*)both.data(), half.size(), 0, t, w1);
        raw::compress2((const uint8_t*)both.data(), both.size(), half.size(), t, w2);
        const std::string c1 = w1.str(), c2 = w2.str();
        CHECK(c1.size() > half.size() && c2.size() < half.size() / 20);
        std::vector<uint8_t> first, second;
        raw::decompress_raw((const uint8_t*)c1.data(), c1.size(), nullptr, 0, first, size_t(1) << 30);
        CHECK(std::string(first.begin(), first.end()) == half);
        raw::decompress_raw((const uint8_t*)c2.data(), c2.size(), (const uint8_t*)both.data(), half.size(), second, size_t(1) << 30);
        CHECK(std::string(second.begin(), second.end()) == half);
        bool threw = false;                                       // without the history its matches point nowhere
        try { std::vector<uint8_t> o3; raw::decompress_raw((const uint8_t*)c2.data(), c2.size(), nullptr, 0, o3, size_t(1) << 30); }
        catch (const raw::DecodeError& e) { threw = e.kind == raw::DecodeError::InvalidDeduplicationOffset; }
        CHECK(threw);
    }
    {   // frames: CompressionSettings -> LZ4FrameReader, block by block and all at once
        std::string big;
        for (int i = 0; i < 30000; i++) big += "lorem ipsum " + std::to_string(i * 7919 % 1000) + " ";
        std::istringstream in(big);
        std::ostringstream frame;
        framed::CompressionSettings().block_size(64 << 10).block_checksums(true).compress(in, frame);
        const std::string f = frame.str();
        CHECK(f.size() < big.size() && (uint8_t)f[0] == 0x04 && (uint8_t)f[3] == 0x18);
        std::istringstream r1(f);
        framed::LZ4FrameReader reader(r1);
        CHECK(reader.block_size() == (64 << 10));
        std::vector<uint8_t> plain;
        reader.into_read().read_to_end(plain);
        CHECK(std::string(plain.begin(), plain.end()) == big);
        std::istringstream r2(f);
        const std::vector<uint8_t> all = framed::decompress_frame(r2);
        CHECK(std::string(all.begin(), all.end()) == big);
        std::string bad = f;
        bad[bad.size() - 1] ^= 1;
        std::istringstream r3(bad);
        bool threw = false;
        try { framed::decompress_frame(r3); } catch (const framed::FrameError& e) { threw = e.status == LZF_F_FRAME_CHECKSUM_FAIL; }
        CHECK(threw);
        threw = false;
        try { std::istringstream i2(big); std::ostringstream o2; framed::CompressionSettings().block_size(12345).compress(i2, o2); }
        catch (const framed::FrameError& e) { threw = e.status == LZF_F_INVALID_BLOCK_SIZE; }
        CHECK(threw);
    }
    std::printf(fails ? "HOST-MIRROR-FAILED\n" : "HOST-MIRROR-OK\n");
    return fails ? 1 : 0;
}
