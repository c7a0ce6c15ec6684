/* xbyak/xbyak.h -- NOT Xbyak.  TEST INFRASTRUCTURE (oracle/_ref), never linked into the product.
 *
 * A small, independent stand-in for the subset of the Xbyak JIT assembler API
 * (herumi/xbyak, pinned by the reference at fe083912c8ac7b7e2b0081cbd6213997bc8b56e6,
 * cmake/external/xbyak.cmake:31-32 -- not on disk, no network) that deep-fusion's kernel generators
 * use (src/jit_generator.h, src/jit_conv_kernel.cc, src/jit_concat_kernel.cc).  It exists so that
 * the reference's UNMODIFIED generator sources can be compiled where they lie and *executed*:
 *
 *   - every mnemonic call RECORDS one instruction (opcode + operands) instead of encoding bytes;
 *   - CodeGenerator::getCode() hands out a real function pointer (one of a pool of static
 *     trampolines) that runs the recorded stream through CodeGenerator::run();
 *   - run() is an interpreter: general-purpose register ops are done in C++, every vector
 *     instruction is executed by the SAME machine instruction through its AVX-512 intrinsic
 *     (vmaxps through inline asm so the operand order cannot be commuted), on a 32 x 512-bit
 *     register file kept in memory.  Memory operands touch exactly the bytes the instruction
 *     would touch (masked loads / stores for xmm / ymm forms).
 *
 * So the instruction ORDER, operand choice, addressing and rounding-mode attributes are the
 * reference generator's, and the per-instruction arithmetic is the host CPU's.  Needs a host with
 * AVX-512 F/BW/VL/DQ (VNNI for the vpdpbusd path); oracle/Makefile builds it with those -m flags.
 *
 * Only what the reference emits is implemented; anything else aborts with a message.
 */
#ifndef DF_XBYAK_SHIM_H_
#define DF_XBYAK_SHIM_H_

#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <utility>
#include <vector>

namespace Xbyak {

typedef unsigned char uint8;
typedef unsigned int uint32;
typedef uint64_t uint64;

enum LabelType { T_SHORT, T_NEAR, T_AUTO };

[[noreturn]] inline void shim_die(const char* what) {
  fprintf(stderr, "xbyak shim: %s\n", what);
  abort();
}

/* ------------------------------------------------------------------------------ operands */
class Operand {
 public:
  enum Kind { NONE = 0, MEM = 1, REG = 2, XMM = 16, YMM = 32, ZMM = 64 };
  enum Code { RAX = 0, RCX, RDX, RBX, RSP, RBP, RSI, RDI, R8, R9, R10, R11, R12, R13, R14, R15 };
  Operand() : idx_(0), kind_(NONE), bit_(0), rounding_(0) {}
  Operand(int idx, int kind, int bit) : idx_(idx), kind_(kind), bit_(bit), rounding_(0) {}
  int getIdx() const { return idx_; }
  int getBit() const { return bit_; }
  int getKind() const { return kind_; }
  bool isMEM() const { return kind_ == MEM; }
  bool isREG() const { return kind_ == REG; }
  int getRounding() const { return rounding_; }

 protected:
  int idx_, kind_, bit_;
  int rounding_;  /* EVEX embedded rounding: 0 none, 1 rn-sae, 2 rd-sae, 3 ru-sae, 4 rz-sae */
};

class Reg : public Operand {
 public:
  Reg() {}
  Reg(int idx, int kind, int bit) : Operand(idx, kind, bit) {}
  Reg cvt8() const { return Reg(idx_, REG, 8); }
  class Reg16 cvt16() const;
  class Reg32 cvt32() const;
  class Reg64 cvt64() const;
};
class Reg8 : public Reg {
 public:
  explicit Reg8(int idx = 0) : Reg(idx, REG, 8) {}
};
class Reg16 : public Reg {
 public:
  explicit Reg16(int idx = 0) : Reg(idx, REG, 16) {}
};
class Reg32 : public Reg {
 public:
  explicit Reg32(int idx = 0) : Reg(idx, REG, 32) {}
};
class Reg64 : public Reg {
 public:
  explicit Reg64(int idx = 0) : Reg(idx, REG, 64) {}
};
inline Reg16 Reg::cvt16() const { return Reg16(idx_); }
inline Reg32 Reg::cvt32() const { return Reg32(idx_); }
inline Reg64 Reg::cvt64() const { return Reg64(idx_); }

struct EvexModifierRounding {
  explicit EvexModifierRounding(int r) : rounding(r) {}
  int rounding;
};

class Xmm : public Reg {
 public:
  explicit Xmm(int idx = 0, int kind = XMM, int bit = 128) : Reg(idx, kind, bit) {}
};
class Ymm : public Xmm {
 public:
  explicit Ymm(int idx = 0, int kind = YMM, int bit = 256) : Xmm(idx, kind, bit) {}
};
class Zmm : public Ymm {
 public:
  explicit Zmm(int idx = 0) : Ymm(idx, ZMM, 512) {}
  Zmm operator|(const EvexModifierRounding& emr) const {
    Zmm r(*this);
    r.rounding_ = emr.rounding;
    return r;
  }
};

/* base + index * scale + disp */
class RegExp {
 public:
  RegExp() : base(-1), index(-1), scale(0), disp(0) {}
  RegExp(const Reg& r) : base(r.getIdx()), index(-1), scale(0), disp(0) {  // NOLINT: implicit, like Xbyak
    if (r.getBit() != 64) shim_die("address registers must be 64-bit");
  }
  int base, index, scale;
  long long disp;
};
inline RegExp operator+(const RegExp& a, const RegExp& b) {
  RegExp r = a;
  if (b.base >= 0) {
    if (r.base < 0) r.base = b.base;
    else if (r.index < 0) { r.index = b.base; r.scale = 1; }
    else shim_die("too many registers in an address");
  }
  if (b.index >= 0) {
    if (r.index >= 0) shim_die("two index registers in an address");
    r.index = b.index;
    r.scale = b.scale;
  }
  r.disp += b.disp;
  return r;
}
inline RegExp operator+(const RegExp& a, long long disp) {
  RegExp r = a;
  r.disp += disp;
  return r;
}
inline RegExp operator-(const RegExp& a, long long disp) { return a + (-disp); }
inline RegExp operator*(const Reg64& r, int scale) {
  RegExp e;
  e.index = r.getIdx();
  e.scale = scale;
  return e;
}

class Address : public Operand {
 public:
  Address() : broadcast(false) {}
  Address(int bit, bool bcast, const RegExp& e) : Operand(0, MEM, bit), exp(e), broadcast(bcast) {}
  RegExp exp;
  bool broadcast;
};

class AddressFrame {
 public:
  explicit AddressFrame(int bit, bool bcast = false) : bit_(bit), bcast_(bcast) {}
  Address operator[](const RegExp& e) const { return Address(bit_, bcast_, e); }

 private:
  int bit_;
  bool bcast_;
};

class Label {
 public:
  Label() : id(-1) {}
  mutable int id;  /* assigned by the CodeGenerator on first use */
};

/* ------------------------------------------------------------------------- the recorder */
class CodeGenerator;
namespace shim {
enum Op {
  MOV, ADD, SUB, INC, DEC, CMP, XOR, PUSH, POP, RET, JMP, JE, JL, JG, MOVDQU,
  VMOVUPS, VPXORD, VPMAXSW, VPMAXSB, VMAXPS, VPMOVSXBD, VPMOVZXBD, VCVTDQ2PS, VADDPS, VMULPS, VCVTPS2DQ,
  VPMOVSDB, VPMOVUSDB, VPDPBUSD, VPMADDUBSW, VPMADDWD, VPADDD, VMOVD, VPEXTRD, VPBROADCASTD, VPBROADCASTW
};
struct Opnd {
  int kind;  /* Operand::Kind, or -1 for an immediate */
  int idx, bit, rounding;
  RegExp exp;
  bool broadcast;
  long long imm;
  Opnd() : kind(Operand::NONE), idx(0), bit(0), rounding(0), broadcast(false), imm(0) {}
};
struct Inst {
  Op op;
  Opnd a, b, c;
  int imm8;
  int label;
};
const int kSlots = 256;
typedef void (*Entry)(void*);
inline CodeGenerator** slots() {
  static CodeGenerator* s[kSlots];
  return s;
}
inline std::mutex& slot_mutex() {
  static std::mutex m;
  return m;
}
template <int N>
void trampoline(void* arg);
template <int... I>
inline const Entry* entry_table(std::integer_sequence<int, I...>) {
  static const Entry t[] = {&trampoline<I>...};
  return t;
}
inline Entry entry(int i) { return entry_table(std::make_integer_sequence<int, kSlots>())[i]; }
}  // namespace shim

static const EvexModifierRounding T_rn_sae(1), T_rd_sae(2), T_ru_sae(3), T_rz_sae(4);

class CodeGenerator {
 public:
  explicit CodeGenerator(size_t /*maxSize*/ = 4096, void* /*userPtr*/ = 0)
      : rax(0), rcx(1), rdx(2), rbx(3), rsp(4), rbp(5), rsi(6), rdi(7), r8(8), r9(9), r10(10), r11(11), r12(12), r13(13),
        r14(14), r15(15), eax(0), ecx(1), edx(2), ebx(3), esp(4), ebp(5), esi(6), edi(7), r8d(8), r9d(9), r10d(10),
        r11d(11), r12d(12), r13d(13), r14d(14), r15d(15), ptr(0), byte(8), word(16), dword(32), qword(64), xword(128),
        yword(256), zword(512), ptr_b(0, true), xword_b(128, true), yword_b(256, true), zword_b(512, true), slot_(-1) {}
  virtual ~CodeGenerator() {
    if (slot_ >= 0) {
      std::lock_guard<std::mutex> g(shim::slot_mutex());
      shim::slots()[slot_] = 0;
    }
  }
  CodeGenerator(const CodeGenerator&) = delete;
  CodeGenerator& operator=(const CodeGenerator&) = delete;

  const Reg64 rax, rcx, rdx, rbx, rsp, rbp, rsi, rdi, r8, r9, r10, r11, r12, r13, r14, r15;
  const Reg32 eax, ecx, edx, ebx, esp, ebp, esi, edi, r8d, r9d, r10d, r11d, r12d, r13d, r14d, r15d;
  const AddressFrame ptr, byte, word, dword, qword, xword, yword, zword, ptr_b, xword_b, yword_b, zword_b;

  /* the "code": a function pointer that interprets the recorded stream */
  const uint8* getCode() const {
    if (slot_ < 0) {
      std::lock_guard<std::mutex> g(shim::slot_mutex());
      for (int i = 0; i < shim::kSlots && slot_ < 0; ++i)
        if (!shim::slots()[i]) {
          shim::slots()[i] = const_cast<CodeGenerator*>(this);
          slot_ = i;
        }
      if (slot_ < 0) shim_die("out of trampoline slots (too many live kernels)");
    }
    return reinterpret_cast<const uint8*>(shim::entry(slot_));
  }
  size_t getSize() const { return insts_.size() * 16; }
  size_t instructionCount() const { return insts_.size(); }

  void L(const std::string&) { shim_die("string labels are not supported"); }
  void L(const Label& label) {
    if (label.id < 0) label.id = (int)label_pc_.size(), label_pc_.push_back(-1);
    label_pc_[label.id] = (int)insts_.size();
  }

  /* ---- general-purpose instructions */
  void mov(const Operand& d, const Operand& s) { rec2(shim::MOV, d, s); }
  void mov(const Operand& d, long long imm) { rec_imm(shim::MOV, d, imm); }
  void add(const Operand& d, const Operand& s) { rec2(shim::ADD, d, s); }
  void add(const Operand& d, long long imm) { rec_imm(shim::ADD, d, imm); }
  void sub(const Operand& d, const Operand& s) { rec2(shim::SUB, d, s); }
  void sub(const Operand& d, long long imm) { rec_imm(shim::SUB, d, imm); }
  void cmp(const Operand& d, const Operand& s) { rec2(shim::CMP, d, s); }
  void cmp(const Operand& d, long long imm) { rec_imm(shim::CMP, d, imm); }
  void xor_(const Operand& d, const Operand& s) { rec2(shim::XOR, d, s); }
  void inc(const Operand& d) { rec1(shim::INC, d); }
  void dec(const Operand& d) { rec1(shim::DEC, d); }
  void push(const Operand& d) { rec1(shim::PUSH, d); }
  void pop(const Operand& d) { rec1(shim::POP, d); }
  void ret() {
    shim::Inst i = blank(shim::RET);
    insts_.push_back(i);
  }
  void jmp(const Label& l, LabelType = T_AUTO) { rec_jump(shim::JMP, l); }
  void je(const Label& l, LabelType = T_AUTO) { rec_jump(shim::JE, l); }
  void jl(const Label& l, LabelType = T_AUTO) { rec_jump(shim::JL, l); }
  void jg(const Label& l, LabelType = T_AUTO) { rec_jump(shim::JG, l); }
  void movdqu(const Operand& d, const Operand& s) { rec2(shim::MOVDQU, d, s); }

  /* ---- vector instructions (AVX-512 forms used by the reference) */
  void vmovups(const Operand& d, const Operand& s) { rec2(shim::VMOVUPS, d, s); }
  void vpxord(const Xmm& d, const Xmm& a, const Operand& b) { rec3(shim::VPXORD, d, a, b); }
  void vpmaxsw(const Xmm& d, const Xmm& a, const Operand& b) { rec3(shim::VPMAXSW, d, a, b); }
  void vpmaxsb(const Xmm& d, const Xmm& a, const Operand& b) { rec3(shim::VPMAXSB, d, a, b); }
  void vmaxps(const Xmm& d, const Xmm& a, const Operand& b) { rec3(shim::VMAXPS, d, a, b); }
  void vaddps(const Xmm& d, const Xmm& a, const Operand& b) { rec3(shim::VADDPS, d, a, b); }
  void vmulps(const Xmm& d, const Xmm& a, const Operand& b) { rec3(shim::VMULPS, d, a, b); }
  void vpdpbusd(const Xmm& d, const Xmm& a, const Operand& b) { rec3(shim::VPDPBUSD, d, a, b); }
  void vpmaddubsw(const Xmm& d, const Xmm& a, const Operand& b) { rec3(shim::VPMADDUBSW, d, a, b); }
  void vpmaddwd(const Xmm& d, const Xmm& a, const Operand& b) { rec3(shim::VPMADDWD, d, a, b); }
  void vpaddd(const Xmm& d, const Xmm& a, const Operand& b) { rec3(shim::VPADDD, d, a, b); }
  void vpmovsxbd(const Xmm& d, const Operand& s) { rec2(shim::VPMOVSXBD, d, s); }
  void vpmovzxbd(const Xmm& d, const Operand& s) { rec2(shim::VPMOVZXBD, d, s); }
  void vcvtdq2ps(const Xmm& d, const Operand& s) { rec2(shim::VCVTDQ2PS, d, s); }
  void vcvtps2dq(const Xmm& d, const Operand& s) { rec2(shim::VCVTPS2DQ, d, s); }
  void vpmovsdb(const Operand& d, const Xmm& s) { rec2(shim::VPMOVSDB, d, s); }
  void vpmovusdb(const Operand& d, const Xmm& s) { rec2(shim::VPMOVUSDB, d, s); }
  void vmovd(const Operand& d, const Operand& s) { rec2(shim::VMOVD, d, s); }
  void vpextrd(const Operand& d, const Xmm& s, int imm) {
    rec2(shim::VPEXTRD, d, s);
    insts_.back().imm8 = imm;
  }
  void vpbroadcastd(const Xmm& d, const Operand& s) { rec2(shim::VPBROADCASTD, d, s); }
  void vpbroadcastw(const Xmm& d, const Operand& s) { rec2(shim::VPBROADCASTW, d, s); }

  /* ---- the interpreter: System V entry, first argument in rdi */
  void run(void* arg) const;

 private:
  static shim::Opnd conv(const Operand& o) {
    shim::Opnd r;
    r.kind = o.getKind();
    r.idx = o.getIdx();
    r.bit = o.getBit();
    r.rounding = o.getRounding();
    if (o.isMEM()) {
      const Address& a = static_cast<const Address&>(o);
      r.exp = a.exp;
      r.broadcast = a.broadcast;
    }
    return r;
  }
  static shim::Inst blank(shim::Op op) {
    shim::Inst i;
    i.op = op;
    i.imm8 = 0;
    i.label = -1;
    return i;
  }
  void rec1(shim::Op op, const Operand& a) {
    shim::Inst i = blank(op);
    i.a = conv(a);
    insts_.push_back(i);
  }
  void rec2(shim::Op op, const Operand& a, const Operand& b) {
    shim::Inst i = blank(op);
    i.a = conv(a);
    i.b = conv(b);
    insts_.push_back(i);
  }
  void rec3(shim::Op op, const Operand& a, const Operand& b, const Operand& c) {
    shim::Inst i = blank(op);
    i.a = conv(a);
    i.b = conv(b);
    i.c = conv(c);
    insts_.push_back(i);
  }
  void rec_imm(shim::Op op, const Operand& a, long long imm) {
    shim::Inst i = blank(op);
    i.a = conv(a);
    i.b.kind = -1;
    i.b.imm = imm;
    insts_.push_back(i);
  }
  void rec_jump(shim::Op op, const Label& l) {
    if (l.id < 0) l.id = (int)label_pc_.size(), label_pc_.push_back(-1);
    shim::Inst i = blank(op);
    i.label = l.id;
    insts_.push_back(i);
  }

  std::vector<shim::Inst> insts_;
  std::vector<int> label_pc_;
  mutable int slot_;
};

namespace shim {
template <int N>
void trampoline(void* arg) {
  slots()[N]->run(arg);
}

struct Machine {
  uint64_t gpr[16];
  alignas(64) __m512i z[32];
  long long cmp_a, cmp_b;
  alignas(16) uint64_t stack[128];
};

inline uint64_t ea(const Machine& m, const Opnd& o) {
  uint64_t a = (uint64_t)o.exp.disp;
  if (o.exp.base >= 0) a += m.gpr[o.exp.base];
  if (o.exp.index >= 0) a += m.gpr[o.exp.index] * (uint64_t)o.exp.scale;
  return a;
}
inline uint64_t read_gpr(const Machine& m, const Opnd& o) {
  const uint64_t v = m.gpr[o.idx];
  return o.bit == 64 ? v : (o.bit == 32 ? (v & 0xffffffffull) : (o.bit == 16 ? (v & 0xffffull) : (v & 0xffull)));
}
inline void write_gpr(Machine& m, const Opnd& o, uint64_t v) {
  if (o.bit == 64) m.gpr[o.idx] = v;
  else if (o.bit == 32) m.gpr[o.idx] = v & 0xffffffffull;  /* 32-bit writes zero-extend */
  else if (o.bit == 16) m.gpr[o.idx] = (m.gpr[o.idx] & ~0xffffull) | (v & 0xffffull);
  else m.gpr[o.idx] = (m.gpr[o.idx] & ~0xffull) | (v & 0xffull);
}
inline long long sext(uint64_t v, int bit) {
  return bit == 64 ? (long long)v : (bit == 32 ? (long long)(int32_t)v : (bit == 16 ? (long long)(int16_t)v : (long long)(int8_t)v));
}
inline __mmask16 lanes32(int bit) { return bit == 512 ? 0xFFFF : (bit == 256 ? 0x00FF : 0x000F); }
/* vector source operand: register, or memory of the destination's width (or a 4-byte broadcast) */
inline __m512i vsrc(const Machine& m, const Opnd& o, int bit) {
  if (o.kind != Operand::MEM) return m.z[o.idx];
  const void* p = reinterpret_cast<const void*>(ea(m, o));
  if (o.broadcast) {
    int32_t w;
    memcpy(&w, p, 4);
    return _mm512_set1_epi32(w);
  }
  return _mm512_maskz_loadu_epi32(lanes32(bit), p);
}
/* write a vector result of `bit` bits: the upper part of the zmm register is zeroed (VEX / EVEX rule) */
inline void vdst(Machine& m, const Opnd& o, __m512i v, int bit) { m.z[o.idx] = _mm512_maskz_mov_epi32(lanes32(bit), v); }
}  // namespace shim

inline void CodeGenerator::run(void* arg) const {
  using namespace shim;
  Machine m;
  memset(m.gpr, 0, sizeof(m.gpr));
  for (int i = 0; i < 32; ++i) m.z[i] = _mm512_setzero_si512();
  m.cmp_a = m.cmp_b = 0;
  m.gpr[Operand::RSP] = reinterpret_cast<uint64_t>(&m.stack[127]);
  m.gpr[Operand::RDI] = reinterpret_cast<uint64_t>(arg);
  const size_t n = insts_.size();
  size_t pc = 0;
  while (pc < n) {
    const Inst& in = insts_[pc++];
    const Opnd &a = in.a, &b = in.b, &c = in.c;
    switch (in.op) {
      case MOV: {
        if (a.kind == Operand::REG) {
          uint64_t v;
          if (b.kind == -1) v = (uint64_t)b.imm;
          else if (b.kind == Operand::REG) v = read_gpr(m, b);
          else {
            v = 0;
            memcpy(&v, reinterpret_cast<const void*>(ea(m, b)), a.bit / 8);
          }
          write_gpr(m, a, v);
        } else if (a.kind == Operand::MEM && b.kind == Operand::REG) {
          const uint64_t v = read_gpr(m, b);
          memcpy(reinterpret_cast<void*>(ea(m, a)), &v, b.bit / 8);
        } else {
          shim_die("mov: unsupported operand form");
        }
        break;
      }
      case ADD:
      case SUB:
      case XOR: {
        if (a.kind != Operand::REG) shim_die("alu: destination must be a register");
        const uint64_t x = read_gpr(m, a);
        const uint64_t y = b.kind == -1 ? (uint64_t)b.imm : (b.kind == Operand::REG ? read_gpr(m, b) : 0);
        if (b.kind == Operand::MEM) shim_die("alu: memory source not supported");
        const uint64_t r = in.op == ADD ? x + y : (in.op == SUB ? x - y : (x ^ y));
        write_gpr(m, a, r);
        m.cmp_a = sext(r, a.bit);
        m.cmp_b = 0;
        break;
      }
      case INC:
      case DEC: {
        const uint64_t r = read_gpr(m, a) + (in.op == INC ? 1 : (uint64_t)-1);
        write_gpr(m, a, r);
        m.cmp_a = sext(r, a.bit);
        m.cmp_b = 0;
        break;
      }
      case CMP: {
        m.cmp_a = sext(read_gpr(m, a), a.bit);
        m.cmp_b = b.kind == -1 ? b.imm : sext(read_gpr(m, b), b.bit);
        break;
      }
      case PUSH:
        m.gpr[Operand::RSP] -= 8;
        memcpy(reinterpret_cast<void*>(m.gpr[Operand::RSP]), &m.gpr[a.idx], 8);
        break;
      case POP:
        memcpy(&m.gpr[a.idx], reinterpret_cast<const void*>(m.gpr[Operand::RSP]), 8);
        m.gpr[Operand::RSP] += 8;
        break;
      case RET: return;
      case JMP: pc = (size_t)label_pc_[in.label]; break;
      case JE: if (m.cmp_a == m.cmp_b) pc = (size_t)label_pc_[in.label]; break;
      case JL: if (m.cmp_a < m.cmp_b) pc = (size_t)label_pc_[in.label]; break;
      case JG: if (m.cmp_a > m.cmp_b) pc = (size_t)label_pc_[in.label]; break;
      case MOVDQU:
        if (a.kind == Operand::MEM) _mm_storeu_si128(reinterpret_cast<__m128i*>(ea(m, a)), _mm512_castsi512_si128(m.z[b.idx]));
        else vdst(m, a, _mm512_castsi128_si512(_mm_loadu_si128(reinterpret_cast<const __m128i*>(ea(m, b)))), 128);
        break;
      case VMOVUPS:
        if (a.kind == Operand::MEM) _mm512_mask_storeu_epi32(reinterpret_cast<void*>(ea(m, a)), lanes32(b.bit), m.z[b.idx]);
        else vdst(m, a, vsrc(m, b, a.bit), a.bit);
        break;
      case VPXORD: vdst(m, a, _mm512_xor_epi32(m.z[b.idx], vsrc(m, c, a.bit)), a.bit); break;
      case VPMAXSW: vdst(m, a, _mm512_max_epi16(m.z[b.idx], vsrc(m, c, a.bit)), a.bit); break;
      case VPMAXSB: vdst(m, a, _mm512_max_epi8(m.z[b.idx], vsrc(m, c, a.bit)), a.bit); break;
      case VPADDD: vdst(m, a, _mm512_add_epi32(m.z[b.idx], vsrc(m, c, a.bit)), a.bit); break;
      case VPMADDUBSW: vdst(m, a, _mm512_maddubs_epi16(m.z[b.idx], vsrc(m, c, a.bit)), a.bit); break;
      case VPMADDWD: vdst(m, a, _mm512_madd_epi16(m.z[b.idx], vsrc(m, c, a.bit)), a.bit); break;
      case VPDPBUSD: vdst(m, a, _mm512_dpbusd_epi32(m.z[a.idx], m.z[b.idx], vsrc(m, c, a.bit)), a.bit); break;
      case VMAXPS: {  /* dst = src1 > src2 ? src1 : src2 (src2 on NaN / equal): keep the operand order */
        const __m512 s1 = _mm512_castsi512_ps(m.z[b.idx]), s2 = _mm512_castsi512_ps(vsrc(m, c, a.bit));
        __m512 r;
        __asm__("vmaxps %2, %1, %0" : "=v"(r) : "v"(s1), "v"(s2));
        vdst(m, a, _mm512_castps_si512(r), a.bit);
        break;
      }
      case VADDPS:
        vdst(m, a, _mm512_castps_si512(_mm512_add_ps(_mm512_castsi512_ps(m.z[b.idx]), _mm512_castsi512_ps(vsrc(m, c, a.bit)))), a.bit);
        break;
      case VMULPS:
        vdst(m, a, _mm512_castps_si512(_mm512_mul_ps(_mm512_castsi512_ps(m.z[b.idx]), _mm512_castsi512_ps(vsrc(m, c, a.bit)))), a.bit);
        break;
      case VCVTDQ2PS: vdst(m, a, _mm512_castps_si512(_mm512_cvtepi32_ps(vsrc(m, b, a.bit))), a.bit); break;
      case VCVTPS2DQ: {
        const __m512 s = _mm512_castsi512_ps(vsrc(m, b, a.bit));
        __m512i r;
        switch (a.rounding) {
          case 0: r = _mm512_cvtps_epi32(s); break;
          case 1: r = _mm512_cvt_roundps_epi32(s, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC); break;
          case 2: r = _mm512_cvt_roundps_epi32(s, _MM_FROUND_TO_NEG_INF | _MM_FROUND_NO_EXC); break;
          case 3: r = _mm512_cvt_roundps_epi32(s, _MM_FROUND_TO_POS_INF | _MM_FROUND_NO_EXC); break;
          default: r = _mm512_cvt_roundps_epi32(s, _MM_FROUND_TO_ZERO | _MM_FROUND_NO_EXC); break;
        }
        vdst(m, a, r, a.bit);
        break;
      }
      case VPMOVSXBD: {
        const __m128i s = b.kind == Operand::MEM ? _mm_loadu_si128(reinterpret_cast<const __m128i*>(ea(m, b))) : _mm512_castsi512_si128(m.z[b.idx]);
        vdst(m, a, _mm512_cvtepi8_epi32(s), a.bit);
        break;
      }
      case VPMOVZXBD: {
        const __m128i s = b.kind == Operand::MEM ? _mm_loadu_si128(reinterpret_cast<const __m128i*>(ea(m, b))) : _mm512_castsi512_si128(m.z[b.idx]);
        vdst(m, a, _mm512_cvtepu8_epi32(s), a.bit);
        break;
      }
      case VPMOVSDB:
      case VPMOVUSDB: {
        const __m128i r = in.op == VPMOVSDB ? _mm512_cvtsepi32_epi8(m.z[b.idx]) : _mm512_cvtusepi32_epi8(m.z[b.idx]);
        if (a.kind == Operand::MEM) _mm_storeu_si128(reinterpret_cast<__m128i*>(ea(m, a)), r);
        else vdst(m, a, _mm512_castsi128_si512(r), 128);
        break;
      }
      case VMOVD:
        if (a.kind == Operand::REG) write_gpr(m, a, (uint32_t)_mm_cvtsi128_si32(_mm512_castsi512_si128(m.z[b.idx])));
        else if (b.kind == Operand::REG) vdst(m, a, _mm512_castsi128_si512(_mm_cvtsi32_si128((int)read_gpr(m, b))), 128);
        else shim_die("vmovd: unsupported operand form");
        break;
      case VPEXTRD: {
        alignas(16) uint32_t w[4];
        _mm_store_si128(reinterpret_cast<__m128i*>(w), _mm512_castsi512_si128(m.z[b.idx]));
        if (a.kind != Operand::REG) shim_die("vpextrd: destination must be a register");
        write_gpr(m, a, w[in.imm8 & 3]);
        break;
      }
      case VPBROADCASTD: {
        int32_t w;
        if (b.kind == Operand::MEM) memcpy(&w, reinterpret_cast<const void*>(ea(m, b)), 4);
        else if (b.kind == Operand::REG) w = (int32_t)read_gpr(m, b);
        else w = _mm_cvtsi128_si32(_mm512_castsi512_si128(m.z[b.idx]));
        vdst(m, a, _mm512_set1_epi32(w), a.bit);
        break;
      }
      case VPBROADCASTW: {
        int16_t w;
        if (b.kind == Operand::MEM) memcpy(&w, reinterpret_cast<const void*>(ea(m, b)), 2);
        else if (b.kind == Operand::REG) w = (int16_t)read_gpr(m, b);
        else w = (int16_t)_mm_cvtsi128_si32(_mm512_castsi512_si128(m.z[b.idx]));
        vdst(m, a, _mm512_set1_epi16(w), a.bit);
        break;
      }
    }
  }
}

}  // namespace Xbyak
#endif
