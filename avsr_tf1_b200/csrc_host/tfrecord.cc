// TFRecord / tf.train.SequenceExample reader, writer and padded-batch assembler (include/avsr_io.h).
// Host-only C++17; no TensorFlow, no protobuf library: the two wire formats are small enough to speak directly.
//
// Reference behaviour restated here (file:line into the reference tree):
//   framing                       tf.data.TFRecordDataset / tf.python_io (io_utils.py:93,97,309)
//   SequenceExample schema        dataset_writer.py:290-311 (labels), :439-458 (features), :461-498 (video [+ aus])
//   stream inspection             io_utils.py:308-341 _get_input_shape_from_record
//   example parsing               io_utils.py:21-86 _parse_input_function / _parse_labels_function (EOS appended :80-83)
//   padded_batch                  io_utils.py:109-122
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/avsr_io.h"

namespace {

// Error text of the last failing call of THIS thread.  Worker threads of a call record into their own thread-local
// buffer; the calling thread collects the first worker message under a mutex (collect_worker_error) so that
// avsr_io_last_error never returns a stale message of an earlier call or a torn one.
thread_local char g_err[512] = "";
std::mutex g_err_mutex;
int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
void clear_error() { g_err[0] = 0; }
// called by a worker thread after a failing step: keeps the first message of the call in `first`
void collect_worker_error(char (&first)[512]) {
  std::lock_guard<std::mutex> lock(g_err_mutex);
  if (!first[0]) memcpy(first, g_err, sizeof(g_err));
}

// ---- crc32c ---------------------------------------------------------------------------------------
uint32_t g_tab[8][256];
bool g_tab_ready = false;
void init_tables() {
  if (g_tab_ready) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
    g_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int s = 1; s < 8; ++s) g_tab[s][i] = (g_tab[s - 1][i] >> 8) ^ g_tab[0][g_tab[s - 1][i] & 0xFF];
  g_tab_ready = true;
}
struct TableInit {
  TableInit() { init_tables(); }
} g_table_init;

uint32_t crc32c_sw(uint32_t crc, const uint8_t* p, size_t n) {
  while (n && ((uintptr_t)p & 7)) {
    crc = (crc >> 8) ^ g_tab[0][(crc ^ *p++) & 0xFF];
    --n;
  }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= crc;
    crc = g_tab[7][w & 0xFF] ^ g_tab[6][(w >> 8) & 0xFF] ^ g_tab[5][(w >> 16) & 0xFF] ^ g_tab[4][(w >> 24) & 0xFF] ^
          g_tab[3][(w >> 32) & 0xFF] ^ g_tab[2][(w >> 40) & 0xFF] ^ g_tab[1][(w >> 48) & 0xFF] ^ g_tab[0][w >> 56];
    p += 8;
    n -= 8;
  }
  while (n--) crc = (crc >> 8) ^ g_tab[0][(crc ^ *p++) & 0xFF];
  return crc;
}
#if defined(__x86_64__)
__attribute__((target("sse4.2"))) uint32_t crc32c_hw(uint32_t crc, const uint8_t* p, size_t n) {
  uint64_t c = crc;
  while (n && ((uintptr_t)p & 7)) {
    c = __builtin_ia32_crc32qi((uint32_t)c, *p++);
    --n;
  }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    c = __builtin_ia32_crc32di(c, w);
    p += 8;
    n -= 8;
  }
  while (n--) c = __builtin_ia32_crc32qi((uint32_t)c, *p++);
  return (uint32_t)c;
}
bool have_hw() {
  static const bool v = __builtin_cpu_supports("sse4.2");
  return v;
}
#else
uint32_t crc32c_hw(uint32_t crc, const uint8_t* p, size_t n) { return crc32c_sw(crc, p, n); }
bool have_hw() { return false; }
#endif
uint32_t crc32c(const void* data, size_t n) {
  const uint8_t* p = (const uint8_t*)data;
  const uint32_t c = have_hw() ? crc32c_hw(0xFFFFFFFFu, p, n) : crc32c_sw(0xFFFFFFFFu, p, n);
  return c ^ 0xFFFFFFFFu;
}
uint32_t mask_crc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xA282EAD8u; }

// ---- protobuf wire format ---------------------------------------------------------------------------
struct Span {
  const uint8_t* p;
  const uint8_t* e;
  bool empty() const { return p >= e; }
  size_t size() const { return (size_t)(e - p); }
};
bool varint(Span& s, uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; shift < 64 && s.p < s.e; shift += 7) {
    const uint8_t b = *s.p++;
    r |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      *v = r;
      return true;
    }
  }
  return false;
}
// next field of a message: number, wire type; for length-delimited fields `sub` is the payload
bool next_field(Span& s, uint32_t* num, uint32_t* wt, Span* sub, uint64_t* val) {
  uint64_t key;
  if (!varint(s, &key)) return false;
  *num = (uint32_t)(key >> 3);
  *wt = (uint32_t)(key & 7);
  switch (*wt) {
    case 0:
      return varint(s, val);
    case 1:
      if (s.size() < 8) return false;
      memcpy(val, s.p, 8);
      s.p += 8;
      return true;
    case 2: {
      uint64_t n;
      if (!varint(s, &n) || n > s.size()) return false;
      sub->p = s.p;
      sub->e = s.p + n;
      s.p += n;
      return true;
    }
    case 5: {
      if (s.size() < 4) return false;
      uint32_t v32;
      memcpy(&v32, s.p, 4);
      *val = v32;
      s.p += 4;
      return true;
    }
    default:
      return false;
  }
}
bool key_is(const Span& k, const char* name) {
  const size_t n = strlen(name);
  return k.size() == n && memcmp(k.p, name, n) == 0;
}
// map<string, X> entry {1: key, 2: value}
bool map_entry(Span entry, Span* key, Span* value) {
  key->p = key->e = value->p = value->e = nullptr;
  uint32_t num, wt;
  Span sub;
  uint64_t v;
  while (!entry.empty()) {
    if (!next_field(entry, &num, &wt, &sub, &v)) return false;
    if (wt == 2 && num == 1) *key = sub;
    if (wt == 2 && num == 2) *value = sub;
  }
  return key->p != nullptr;
}
// Feature{1: BytesList, 2: FloatList, 3: Int64List}; each list: repeated value = 1 (packed or not)
bool feature_int64(Span feat, int64_t* out) {
  uint32_t num, wt;
  Span sub;
  uint64_t v;
  while (!feat.empty()) {
    if (!next_field(feat, &num, &wt, &sub, &v)) return false;
    if (num == 3 && wt == 2) {
      Span list = sub;
      while (!list.empty()) {
        if (!next_field(list, &num, &wt, &sub, &v)) return false;
        if (num != 1) continue;
        if (wt == 0) {
          *out = (int64_t)v;
          return true;
        }
        if (wt == 2) {  // packed
          uint64_t x;
          if (!varint(sub, &x)) return false;
          *out = (int64_t)x;
          return true;
        }
      }
    }
  }
  return false;
}
bool feature_bytes(Span feat, Span* out) {
  uint32_t num, wt;
  Span sub;
  uint64_t v;
  while (!feat.empty()) {
    if (!next_field(feat, &num, &wt, &sub, &v)) return false;
    if (num == 1 && wt == 2) {
      Span list = sub;
      while (!list.empty()) {
        if (!next_field(list, &num, &wt, &sub, &v)) return false;
        if (num == 1 && wt == 2) {
          *out = sub;
          return true;
        }
      }
    }
  }
  return false;
}
// copies up to cap floats of a Feature's FloatList into dst; returns the number stored in the list (or -1)
long feature_floats(Span feat, float* dst, long cap) {
  uint32_t num, wt;
  Span sub;
  uint64_t v;
  long n = 0;
  while (!feat.empty()) {
    if (!next_field(feat, &num, &wt, &sub, &v)) return -1;
    if (num == 2 && wt == 2) {
      Span list = sub;
      while (!list.empty()) {
        if (!next_field(list, &num, &wt, &sub, &v)) return -1;
        if (num != 1) continue;
        if (wt == 2) {  // packed: little-endian floats
          const long k = (long)(sub.size() / 4);
          const long c = std::max(0l, std::min(k, cap - n));
          if (dst && c > 0) memcpy(dst + n, sub.p, (size_t)c * 4);
          n += k;
        } else if (wt == 5) {
          if (dst && n < cap) {
            const uint32_t bits = (uint32_t)v;
            memcpy(dst + n, &bits, 4);
          }
          ++n;
        }
      }
    }
  }
  return n;
}

struct Example {
  Span context{nullptr, nullptr}, lists{nullptr, nullptr};
};
bool split_example(Span rec, Example* ex) {
  uint32_t num, wt;
  Span sub;
  uint64_t v;
  while (!rec.empty()) {
    if (!next_field(rec, &num, &wt, &sub, &v)) return false;
    if (wt == 2 && num == 1) ex->context = sub;
    if (wt == 2 && num == 2) ex->lists = sub;
  }
  return true;
}
// value Feature of context[name]
bool context_get(Span ctx, const char* name, Span* feat) {
  uint32_t num, wt;
  Span sub, key, val;
  uint64_t v;
  while (!ctx.empty()) {
    if (!next_field(ctx, &num, &wt, &sub, &v)) return false;
    if (num == 1 && wt == 2 && map_entry(sub, &key, &val) && key_is(key, name)) {
      *feat = val;
      return true;
    }
  }
  return false;
}
bool lists_get(Span lists, const char* name, Span* flist) {
  return context_get(lists, name, flist);  // same map<string, .> layout
}

// ---- writer side -----------------------------------------------------------------------------------
void put_varint(std::string& s, uint64_t v) {
  while (v >= 0x80) {
    s.push_back((char)(v | 0x80));
    v >>= 7;
  }
  s.push_back((char)v);
}
void put_len(std::string& s, uint32_t field, const std::string& payload) {
  put_varint(s, (field << 3) | 2);
  put_varint(s, payload.size());
  s += payload;
}
std::string feat_int64(int64_t v) {
  std::string packed, list, feat;
  put_varint(packed, (uint64_t)v);
  put_len(list, 1, packed);
  put_len(feat, 3, list);
  return feat;
}
std::string feat_bytes(const char* b) {
  std::string list, feat;
  put_len(list, 1, std::string(b));
  put_len(feat, 1, list);
  return feat;
}
std::string feat_floats(const float* x, long n) {
  std::string list, feat;
  if (n > 0) put_len(list, 1, std::string((const char*)x, (size_t)n * 4));
  put_len(feat, 2, list);
  return feat;
}
std::string entry(const char* key, const std::string& value) {
  std::string e;
  put_len(e, 1, std::string(key));
  put_len(e, 2, value);
  return e;
}

}  // namespace

struct AvsrIoFile {
  int fd = -1;
  const uint8_t* base = nullptr;
  size_t size = 0;
  std::vector<std::pair<size_t, size_t>> rec;  // payload offset, length
  std::vector<long long> length;               // steps per record
  AvsrIoInfo info{};
  Span record(long long i) const { return Span{base + rec[i].first, base + rec[i].first + rec[i].second}; }
};

struct AvsrIoWriter {
  FILE* f = nullptr;
};

extern "C" {

const char* avsr_io_last_error(void) { return g_err; }
uint32_t avsr_io_crc32c(const void* data, size_t n) { return crc32c(data, n); }
uint32_t avsr_io_masked_crc32c(const void* data, size_t n) { return mask_crc(crc32c(data, n)); }

int avsr_io_open(const char* path, int verify_data, AvsrIoFile** out) {
  clear_error();
  *out = nullptr;
  AvsrIoFile* f = new AvsrIoFile();
  f->fd = open(path, O_RDONLY);
  if (f->fd < 0) {
    delete f;
    return fail("cannot open %s", path);
  }
  struct stat st;
  fstat(f->fd, &st);
  f->size = (size_t)st.st_size;
  if (f->size > 0) {
    void* m = mmap(nullptr, f->size, PROT_READ, MAP_PRIVATE, f->fd, 0);
    if (m == MAP_FAILED) {
      close(f->fd);
      delete f;
      return fail("mmap failed for %s", path);
    }
    f->base = (const uint8_t*)m;
  }
  auto bail = [&](int rc) {
    avsr_io_close(f);
    return rc;
  };
  size_t off = 0;
  while (off < f->size) {
    if (f->size - off < 12) return bail(fail("%s: truncated record header at byte %zu", path, off));
    uint64_t len;
    uint32_t crc;
    memcpy(&len, f->base + off, 8);
    memcpy(&crc, f->base + off + 8, 4);
    if (mask_crc(crc32c(f->base + off, 8)) != crc)
      return bail(fail("%s: corrupted record length at byte %zu (crc mismatch)", path, off));
    if (len > f->size - off - 12 || f->size - off - 12 - len < 4)
      return bail(fail("%s: truncated record at byte %zu", path, off));
    if (verify_data) {
      uint32_t dcrc;
      memcpy(&dcrc, f->base + off + 12 + len, 4);
      if (mask_crc(crc32c(f->base + off + 12, len)) != dcrc)
        return bail(fail("%s: corrupted record data at byte %zu (crc mismatch)", path, off));
    }
    f->rec.emplace_back(off + 12, (size_t)len);
    off += 12 + len + 4;
  }
  AvsrIoInfo& in = f->info;
  in.n_records = (long long)f->rec.size();
  in.feat = 0;
  in.channels = 1;
  f->length.resize(f->rec.size());
  for (size_t i = 0; i < f->rec.size(); ++i) {
    Example ex;
    if (!split_example(f->record((long long)i), &ex)) return bail(fail("%s: record %zu is not a SequenceExample", path, i));
    Span feat;
    int64_t v = 0;
    if (i == 0) {  // io_utils.py:308-341
      if (context_get(ex.context, "labels_length", &feat)) {
        in.kind = AVSR_IO_LABELS;
        in.feat = 1;
        Span u;
        if (context_get(ex.context, "unit", &feat) && feature_bytes(feat, &u)) {
          const size_t n = std::min(u.size(), sizeof(in.unit) - 1);
          memcpy(in.unit, u.p, n);
          in.unit[n] = 0;
        }
      } else if (context_get(ex.context, "input_size", &feat) && feature_int64(feat, &v)) {
        in.kind = AVSR_IO_FEATURE;
        in.feat = v;
      } else {
        int64_t w = 0, h = 0, c = 1;
        if (!(context_get(ex.context, "width", &feat) && feature_int64(feat, &w) &&
              context_get(ex.context, "height", &feat) && feature_int64(feat, &h)))
          return bail(fail("%s: first example has neither input_size nor width/height nor labels_length", path));
        if (context_get(ex.context, "channels", &feat)) feature_int64(feat, &c);
        in.kind = AVSR_IO_VIDEO;
        in.width = (int)w;
        in.height = (int)h;
        in.channels = (int)c;
        in.feat = w * h * c;
      }
      Span fl;
      in.has_aus = lists_get(ex.lists, "aus", &fl) ? 1 : 0;
    }
    const char* key = in.kind == AVSR_IO_LABELS ? "labels_length" : "input_length";
    if (!(context_get(ex.context, key, &feat) && feature_int64(feat, &v)))
      return bail(fail("%s: record %zu has no %s", path, i, key));
    f->length[i] = v;
  }
  *out = f;
  return 0;
}

void avsr_io_close(AvsrIoFile* f) {
  if (!f) return;
  if (f->base) munmap((void*)f->base, f->size);
  if (f->fd >= 0) close(f->fd);
  delete f;
}

int avsr_io_info(const AvsrIoFile* f, AvsrIoInfo* info) {
  *info = f->info;
  return 0;
}

int avsr_io_lengths(AvsrIoFile* f, long long* lengths) {
  std::copy(f->length.begin(), f->length.end(), lengths);
  return 0;
}

int avsr_io_filename(AvsrIoFile* f, long long idx, char* dst, int cap) {
  clear_error();
  if (idx < 0 || idx >= f->info.n_records || cap <= 0) return fail("filename: record %lld out of range", idx);
  Example ex;
  Span feat, name;
  dst[0] = 0;
  if (!split_example(f->record(idx), &ex) || !context_get(ex.context, "filename", &feat) || !feature_bytes(feat, &name))
    return fail("record %lld has no filename", idx);
  const size_t n = std::min(name.size(), (size_t)cap - 1);
  memcpy(dst, name.p, n);
  dst[n] = 0;
  return 0;
}

static int fill_one_input(const AvsrIoFile* f, long long idx, int t_pad, float* row, float* aus_row, int32_t* len,
                          int reverse) {
  const long feat = (long)f->info.feat;
  Example ex;
  Span inputs;
  if (!split_example(f->record(idx), &ex) || !lists_get(ex.lists, "inputs", &inputs))
    return fail("record %lld has no `inputs` feature list", idx);
  const long long T = f->length[idx];
  if (T > t_pad) return fail("record %lld has %lld steps, the batch is padded to %d", idx, T, t_pad);
  auto walk = [&](Span list, float* dst, long width, const char* what) -> int {
    uint32_t num, wt;
    Span sub;
    uint64_t v;
    long long t = 0;
    while (!list.empty()) {
      if (!next_field(list, &num, &wt, &sub, &v)) return fail("record %lld: malformed %s list", idx, what);
      if (num != 1 || wt != 2) continue;
      if (t >= T) return fail("record %lld: more %s steps than input_length = %lld", idx, what, T);
      const long long tt = reverse ? T - 1 - t : t;
      const long n = feature_floats(sub, dst + (size_t)tt * width, width);
      if (n != width) return fail("record %lld: step %lld of %s has %ld values, expected %ld", idx, t, what, n, width);
      ++t;
    }
    if (t != T) return fail("record %lld: %lld %s steps, input_length says %lld", idx, t, what, T);
    return 0;
  };
  if (int rc = walk(inputs, row, feat, "inputs")) return rc;
  memset(row + (size_t)T * feat, 0, (size_t)(t_pad - T) * feat * sizeof(float));
  if (aus_row) {
    Span aus;
    if (!lists_get(ex.lists, "aus", &aus)) return fail("record %lld has no `aus` feature list", idx);
    if (int rc = walk(aus, aus_row, 2, "aus")) return rc;
    memset(aus_row + (size_t)T * 2, 0, (size_t)(t_pad - T) * 2 * sizeof(float));
  }
  *len = (int32_t)T;
  return 0;
}

int avsr_io_fill_inputs(AvsrIoFile* f, const long long* idx, int n, int t_pad, float* dst, float* aus_dst,
                        int32_t* lens, int reverse, int n_threads) {
  clear_error();
  if (f->info.kind == AVSR_IO_LABELS) return fail("fill_inputs on a label record");
  if (aus_dst && !f->info.has_aus) return fail("this record has no Action Units");
  for (int i = 0; i < n; ++i)
    if (idx[i] < 0 || idx[i] >= f->info.n_records) return fail("fill_inputs: record %lld out of range", idx[i]);
  const size_t feat = (size_t)f->info.feat;
  std::atomic<int> next(0), failed(0);
  char first_error[512] = "";
  auto work = [&]() {
    for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
      if (fill_one_input(f, idx[i], t_pad, dst + (size_t)i * t_pad * feat,
                         aus_dst ? aus_dst + (size_t)i * t_pad * 2 : nullptr, lens + i, reverse)) {
        failed.store(1);
        collect_worker_error(first_error);  // the message lives in the worker's thread-local buffer
      }
    }
  };
  const int nt = std::max(1, std::min(n_threads, n));
  if (nt == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int k = 0; k < nt; ++k) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  if (failed.load()) {
    memcpy(g_err, first_error, sizeof(g_err));  // into the CALLER's buffer, which avsr_io_last_error reads
    return 1;
  }
  return 0;
}

int avsr_io_fill_labels(AvsrIoFile* f, const long long* idx, int n, int l_pad, int32_t eos, int32_t* dst,
                        int32_t* lens) {
  clear_error();
  if (f->info.kind != AVSR_IO_LABELS) return fail("fill_labels on an input record");
  for (int i = 0; i < n; ++i) {
    if (idx[i] < 0 || idx[i] >= f->info.n_records) return fail("fill_labels: record %lld out of range", idx[i]);
    Example ex;
    Span labels;
    if (!split_example(f->record(idx[i]), &ex) || !lists_get(ex.lists, "labels", &labels))
      return fail("record %lld has no `labels` feature list", idx[i]);
    int32_t* row = dst + (size_t)i * l_pad;
    memset(row, 0, (size_t)l_pad * sizeof(int32_t));
    uint32_t num, wt;
    Span sub;
    uint64_t v;
    int k = 0;
    while (!labels.empty()) {
      if (!next_field(labels, &num, &wt, &sub, &v)) return fail("record %lld: malformed labels", idx[i]);
      if (num != 1 || wt != 2) continue;
      int64_t id;
      if (!feature_int64(sub, &id)) return fail("record %lld: label %d is not an int64", idx[i], k);
      if (k + 1 >= l_pad) return fail("record %lld: more than %d labels (+EOS)", idx[i], l_pad - 1);
      row[k++] = (int32_t)id;
    }
    if (k != f->length[idx[i]]) return fail("record %lld: %d labels, labels_length says %lld", idx[i], k, f->length[idx[i]]);
    row[k] = eos;  // io_utils.py:80-83
    lens[i] = k + 1;
  }
  return 0;
}

// ---- writer ------------------------------------------------------------------------------------------
int avsr_io_writer_open(const char* path, AvsrIoWriter** out) {
  clear_error();
  *out = nullptr;
  FILE* fp = fopen(path, "wb");
  if (!fp) return fail("cannot create %s", path);
  AvsrIoWriter* w = new AvsrIoWriter();
  w->f = fp;
  *out = w;
  return 0;
}

int avsr_io_writer_close(AvsrIoWriter* w) {
  if (!w) return 0;
  const int rc = fclose(w->f);
  delete w;
  return rc ? fail("close failed") : 0;
}

static int write_record(AvsrIoWriter* w, const std::string& context, const std::string& lists) {
  std::string ex;
  put_len(ex, 1, context);
  put_len(ex, 2, lists);
  const uint64_t len = ex.size();
  const uint32_t c1 = mask_crc(crc32c(&len, 8)), c2 = mask_crc(crc32c(ex.data(), ex.size()));
  if (fwrite(&len, 8, 1, w->f) != 1 || fwrite(&c1, 4, 1, w->f) != 1 ||
      (len && fwrite(ex.data(), ex.size(), 1, w->f) != 1) || fwrite(&c2, 4, 1, w->f) != 1)
    return fail("short write");
  return 0;
}

static std::string float_list(const float* x, int steps, long width) {
  std::string fl;
  for (int t = 0; t < steps; ++t) put_len(fl, 1, feat_floats(x + (size_t)t * width, width));
  return fl;
}

int avsr_io_write_feature(AvsrIoWriter* w, const char* sentence_id, const float* inputs, int steps, int size) {
  std::string ctx, lists;
  put_len(ctx, 1, entry("input_length", feat_int64(steps)));
  put_len(ctx, 1, entry("input_size", feat_int64(size)));
  put_len(ctx, 1, entry("filename", feat_bytes(sentence_id)));
  put_len(lists, 1, entry("inputs", float_list(inputs, steps, size)));
  return write_record(w, ctx, lists);
}

int avsr_io_write_video(AvsrIoWriter* w, const char* sentence_id, const float* frames, int steps, int height,
                        int width, int channels, const float* aus) {
  std::string ctx, lists;
  put_len(ctx, 1, entry("input_length", feat_int64(steps)));
  put_len(ctx, 1, entry("width", feat_int64(width)));
  put_len(ctx, 1, entry("height", feat_int64(height)));
  put_len(ctx, 1, entry("channels", feat_int64(channels)));
  put_len(ctx, 1, entry("filename", feat_bytes(sentence_id)));
  put_len(lists, 1, entry("inputs", float_list(frames, steps, (long)height * width * channels)));
  if (aus) put_len(lists, 1, entry("aus", float_list(aus, steps, 2)));
  return write_record(w, ctx, lists);
}

int avsr_io_write_labels(AvsrIoWriter* w, const char* label_id, const int64_t* labels, int n, const char* unit) {
  std::string ctx, lists, fl;
  put_len(ctx, 1, entry("unit", feat_bytes(unit)));
  put_len(ctx, 1, entry("labels_length", feat_int64(n)));
  put_len(ctx, 1, entry("filename", feat_bytes(label_id)));
  for (int i = 0; i < n; ++i) put_len(fl, 1, feat_int64(labels[i]));
  put_len(lists, 1, entry("labels", fl));
  return write_record(w, ctx, lists);
}

}  // extern "C"
