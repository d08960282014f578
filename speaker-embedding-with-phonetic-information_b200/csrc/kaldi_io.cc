// See kaldi_io.h.  Formats restate upstream Kaldi's src/base/io-funcs.cc,
// src/matrix/kaldi-{matrix,vector}.cc, src/matrix/compressed-matrix.cc,
// src/util/{kaldi-io,kaldi-table-inl,parse-options}.* (not shipped by the
// reference).
#include "kaldi_io.h"

#include <sys/wait.h>
#include <unistd.h>

#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <fstream>
#include <mutex>
#include <thread>
#include <iostream>

namespace cvx {

static std::string g_program = "cvx";

void SetProgramName(const std::string& name) {
  size_t p = name.find_last_of('/');
  g_program = (p == std::string::npos) ? name : name.substr(p + 1);
}

void LogMessage(const char* severity, const char* func, const char* file, int line,
                const std::string& msg) {
  const char* base = std::strrchr(file, '/');
  base = base ? base + 1 : file;
  std::fprintf(stderr, "%s (%s:%s():%s:%d) %s\n", severity, g_program.c_str(), func, base, line,
               msg.c_str());
}

// ---------------------------------------------------------------- Input ----

namespace {

bool PreadAll(int fd, unsigned char* dst, size_t len, off_t at) {
  while (len) {
    ssize_t k = ::pread(fd, dst, len, at);
    if (k <= 0) return false;
    dst += k;
    at += k;
    len -= static_cast<size_t>(k);
  }
  return true;
}

// Helper threads of a reader thread (thread_local): each reads a later part of a large matrix body
// while the owner reads the first.  A persistent thread, because a table has thousands of entries and
// starting a thread per entry costs a fifth of the read itself.
class PreadHelper {
 public:
  ~PreadHelper() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
  }
  void Start(int fd, unsigned char* dst, size_t len, off_t at) {
    if (!th_.joinable()) th_ = std::thread([this] { Run(); });
    std::lock_guard<std::mutex> lk(mu_);
    fd_ = fd, dst_ = dst, len_ = len, at_ = at;
    pending_ = true;
    done_ = false;
    cv_.notify_all();
  }
  bool Wait() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return done_; });
    return ok_;
  }

 private:
  void Run() {
    std::unique_lock<std::mutex> lk(mu_);
    while (true) {
      cv_.wait(lk, [&] { return pending_ || stop_; });
      if (stop_) return;
      pending_ = false;
      const int fd = fd_;
      unsigned char* dst = dst_;
      const size_t len = len_;
      const off_t at = at_;
      lk.unlock();
      const bool ok = PreadAll(fd, dst, len, at);
      lk.lock();
      ok_ = ok;
      done_ = true;
      cv_.notify_all();
    }
  }
  std::thread th_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool pending_ = false, done_ = false, ok_ = false, stop_ = false;
  int fd_ = -1;
  unsigned char* dst_ = nullptr;
  size_t len_ = 0;
  off_t at_ = 0;
};

constexpr int kPreadHelpers = 3;
PreadHelper* ThreadPreadHelpers() {
  thread_local PreadHelper helpers[kPreadHelpers];
  return helpers;
}

}  // namespace

static std::string Trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
  while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
  return s.substr(a, b - a);
}

void Input::Open(const std::string& rx_in) {
  Close();
  std::string rx = Trim(rx_in);
  name_ = rx;
  buf_.resize(1 << 16);  // headers and text go through this buffer; large binary bodies are read straight into their destination
  pos_ = end_ = 0;
  if (rx.empty()) throw KaldiError("empty rxfilename");
  if (rx == "-") {
    fp_ = stdin;
    is_stdin_ = true;
    return;
  }
  if (rx.back() == '|') {
    std::string cmd = Trim(rx.substr(0, rx.size() - 1));
    fp_ = popen(cmd.c_str(), "r");
    if (!fp_) throw KaldiError("cannot run input pipe: " + cmd);
    is_pipe_ = true;
    return;
  }
  // optional ":offset"
  long offset = 0;
  std::string path = rx;
  size_t c = rx.find_last_of(':');
  if (c != std::string::npos && c + 1 < rx.size()) {
    bool digits = true;
    for (size_t i = c + 1; i < rx.size(); ++i) digits = digits && std::isdigit(static_cast<unsigned char>(rx[i]));
    if (digits) {
      offset = std::strtol(rx.c_str() + c + 1, nullptr, 10);
      path = rx.substr(0, c);
    }
  }
  fp_ = std::fopen(path.c_str(), "rb");
  if (!fp_) throw KaldiError("cannot open " + path + ": " + std::strerror(errno));
  if (offset && std::fseek(fp_, offset, SEEK_SET) != 0) throw KaldiError("cannot seek in " + path);
}

int Input::Close() {
  int status = 0;
  if (fp_) {
    if (is_pipe_) {
      int s = pclose(fp_);
      status = (s == -1) ? -1 : (WIFEXITED(s) ? WEXITSTATUS(s) : 128);
    } else if (!is_stdin_) {
      std::fclose(fp_);
    }
  }
  fp_ = nullptr;
  is_pipe_ = is_stdin_ = false;
  pos_ = end_ = 0;
  return status;
}

bool Input::Fill() {
  if (!fp_) return false;
  if (pos_ < end_) return true;
  size_t n = std::fread(buf_.data(), 1, buf_.size(), fp_);
  pos_ = 0;
  end_ = n;
  return n > 0;
}

int Input::Peek() {
  if (pos_ >= end_ && !Fill()) return -1;
  return buf_[pos_];
}

int Input::Get() {
  if (pos_ >= end_ && !Fill()) return -1;
  return buf_[pos_++];
}

void Input::Read(void* dst, size_t n) {
  unsigned char* d = static_cast<unsigned char*>(dst);
  while (n) {
    if (pos_ >= end_ && n >= (size_t{1} << 18) && fp_ && !is_pipe_ && !is_stdin_) {
      // large body from a file: two to four threads pread() a part each -- one thread copies ~4.5 GB/s out of
      // the page cache, one B200 consumes 7.5 GB/s of features on the c-vector net
      const long off = std::ftell(fp_);
      const int fd = fileno(fp_);
      if (off >= 0 && fd >= 0) {
        // this reader thread's persistent helpers take the later parts, the thread itself the first
        PreadHelper* helpers = ThreadPreadHelpers();
        // (a part of >= 128 KB is >= 30 us of copying against ~10 us to wake a helper)
        const int parts = n >= (size_t{1} << 20) ? kPreadHelpers + 1 : 2;
        const size_t step = (n / parts) & ~size_t{4095};
        for (int p = 1; p < parts; ++p) {
          const size_t begin = step * p, end = p + 1 == parts ? n : step * (p + 1);
          helpers[p - 1].Start(fd, d + begin, end - begin, static_cast<off_t>(off) + static_cast<off_t>(begin));
        }
        bool ok = PreadAll(fd, d, step, off);
        for (int p = 1; p < parts; ++p) ok = helpers[p - 1].Wait() && ok;
        if (!ok) throw KaldiError("unexpected end of input in " + name_);
        if (std::fseek(fp_, off + static_cast<long>(n), SEEK_SET) != 0) throw KaldiError("cannot seek in " + name_);
        return;
      }
    }
    if (pos_ >= end_ && n >= buf_.size() && fp_) {
      // large body: read straight into the destination, bypassing the buffer
      size_t got = std::fread(d, 1, n, fp_);
      if (got == 0) throw KaldiError("unexpected end of input in " + name_);
      d += got;
      n -= got;
      continue;
    }
    if (pos_ >= end_ && !Fill()) throw KaldiError("unexpected end of input in " + name_);
    size_t k = std::min(n, end_ - pos_);
    std::memcpy(d, buf_.data() + pos_, k);
    pos_ += k;
    d += k;
    n -= k;
  }
}

void Input::SkipSpace() {
  int c;
  while ((c = Peek()) >= 0 && std::isspace(c)) ++pos_;
}

std::string Input::ReadToken() {
  SkipSpace();
  std::string t;
  int c;
  while ((c = Peek()) >= 0 && !std::isspace(c)) {
    t.push_back(static_cast<char>(c));
    ++pos_;
  }
  if (c >= 0) ++pos_;  // exactly one whitespace character
  return t;
}

std::string Input::PeekToken() {
  // Only used on text input right after a token boundary: make sure enough
  // bytes are buffered by compacting and refilling.
  if (end_ - pos_ < 256 && fp_) {
    size_t left = end_ - pos_;
    std::memmove(buf_.data(), buf_.data() + pos_, left);
    size_t n = std::fread(buf_.data() + left, 1, buf_.size() - left, fp_);
    pos_ = 0;
    end_ = left + n;
  }
  size_t p = pos_;
  while (p < end_ && std::isspace(buf_[p])) ++p;
  std::string t;
  while (p < end_ && !std::isspace(buf_[p])) t.push_back(static_cast<char>(buf_[p++]));
  return t;
}

std::string Input::ReadLine() {
  std::string s;
  int c;
  while ((c = Get()) >= 0 && c != '\n') s.push_back(static_cast<char>(c));
  if (!s.empty() && s.back() == '\r') s.pop_back();
  return s;
}

// --------------------------------------------------------------- Output ----

void Output::Open(const std::string& wx_in) {
  Close();
  std::string wx = Trim(wx_in);
  name_ = wx;
  bytes_ = 0;
  if (wx.empty() || wx == "-") {
    fp_ = stdout;
    is_stdout_ = true;
    return;
  }
  if (wx[0] == '|') {
    fp_ = popen(Trim(wx.substr(1)).c_str(), "w");
    if (!fp_) throw KaldiError("cannot run output pipe: " + wx);
    is_pipe_ = true;
    return;
  }
  fp_ = std::fopen(wx.c_str(), "wb");
  if (!fp_) throw KaldiError("cannot open " + wx + " for writing: " + std::strerror(errno));
}

int Output::Close() {
  int status = 0;
  if (fp_) {
    if (is_pipe_) {
      int s = pclose(fp_);
      status = (s == -1) ? -1 : (WIFEXITED(s) ? WEXITSTATUS(s) : 128);
    } else if (is_stdout_) {
      std::fflush(fp_);
    } else {
      if (std::fclose(fp_) != 0) status = -1;
    }
  }
  fp_ = nullptr;
  is_pipe_ = is_stdout_ = false;
  return status;
}

void Output::Write(const void* src, size_t n) {
  if (!fp_) throw KaldiError("write to closed output " + name_);
  if (n && std::fwrite(src, 1, n, fp_) != n) throw KaldiError("write failed on " + name_);
  bytes_ += static_cast<int64_t>(n);
}

void Output::Flush() {
  if (fp_) std::fflush(fp_);
}

// ---------------------------------------------------------- basic types ----

bool ReadBinaryHeader(Input& in) {
  if (in.Peek() == '\0') {
    in.Get();
    if (in.Get() != 'B') throw KaldiError("bad binary header in " + in.Name());
    return true;
  }
  return false;
}

void ExpectToken(Input& in, const std::string& tok) {
  std::string t = in.ReadToken();
  if (t != tok) throw KaldiError("expected token " + tok + ", got '" + t + "' in " + in.Name());
}

int32_t ReadInt32(Input& in, bool binary) {
  if (binary) {
    int sz = in.Get();
    if (sz != 4) throw KaldiError("expected 4-byte integer in " + in.Name());
    int32_t v;
    in.Read(&v, 4);
    return v;
  }
  std::string t = in.ReadToken();
  char* e = nullptr;
  long v = std::strtol(t.c_str(), &e, 10);
  if (t.empty() || *e) throw KaldiError("bad integer '" + t + "' in " + in.Name());
  return static_cast<int32_t>(v);
}

static float ParseFloatText(const std::string& t, const std::string& where) {
  if (t == "inf" || t == "Inf" || t == "+inf") return INFINITY;
  if (t == "-inf" || t == "-Inf") return -INFINITY;
  if (t == "nan" || t == "NaN" || t == "-nan") return NAN;
  char* e = nullptr;
  float v = std::strtof(t.c_str(), &e);
  if (t.empty() || *e) throw KaldiError("bad float '" + t + "' in " + where);
  return v;
}

float ReadFloat(Input& in, bool binary) {
  if (binary) {
    int sz = in.Get();
    if (sz == 4) {
      float v;
      in.Read(&v, 4);
      return v;
    }
    if (sz == 8) {
      double v;
      in.Read(&v, 8);
      return static_cast<float>(v);
    }
    throw KaldiError("expected 4- or 8-byte real in " + in.Name());
  }
  return ParseFloatText(in.ReadToken(), in.Name());
}

bool ReadBool(Input& in, bool binary) {
  if (!binary) in.SkipSpace();
  int c = in.Get();
  if (c != 'T' && c != 'F') throw KaldiError("bad bool in " + in.Name());
  if (!binary) {
    int n = in.Peek();
    if (n >= 0 && std::isspace(n)) in.Get();
  }
  return c == 'T';
}

// Text "[ ... ]": rows end at '\n' (or at ']').
static void ReadTextMatrix(Input& in, Matrix* m) {
  in.SkipSpace();
  if (in.Get() != '[') throw KaldiError("expected '[' in " + in.Name());
  m->rows = m->cols = 0;
  m->data.clear();
  int cur_cols = 0;
  std::string tok;
  auto end_row = [&]() {
    if (cur_cols == 0) return;
    if (m->rows == 0) m->cols = cur_cols;
    if (cur_cols != m->cols) throw KaldiError("ragged text matrix in " + in.Name());
    ++m->rows;
    cur_cols = 0;
  };
  while (true) {
    int c = in.Get();
    if (c < 0) throw KaldiError("unterminated text matrix in " + in.Name());
    if (std::isspace(c) || c == ']') {
      if (!tok.empty()) {
        m->data.push_back(ParseFloatText(tok, in.Name()));
        tok.clear();
        ++cur_cols;
      }
      if (c == '\n' || c == ']') end_row();
      if (c == ']') break;
    } else {
      tok.push_back(static_cast<char>(c));
    }
  }
}

static void ReadCompressed(Input& in, const std::string& tag, Matrix* m) {
  float min_value, range;
  int32_t rows, cols;
  in.Read(&min_value, 4);
  in.Read(&range, 4);
  in.Read(&rows, 4);
  in.Read(&cols, 4);
  if (rows < 0 || cols < 0) throw KaldiError("bad compressed matrix header in " + in.Name());
  m->rows = rows;
  m->cols = cols;
  m->data.assign(static_cast<size_t>(rows) * cols, 0.f);
  const size_t n = static_cast<size_t>(rows) * cols;
  if (tag == "CM") {
    std::vector<uint16_t> hdr(static_cast<size_t>(cols) * 4);
    in.Read(hdr.data(), hdr.size() * 2);
    std::vector<uint8_t> bytes(n);
    in.Read(bytes.data(), n);
    const float inc = range * (1.0f / 65535.0f);
    for (int32_t c = 0; c < cols; ++c) {
      float p0 = min_value + inc * hdr[c * 4 + 0], p25 = min_value + inc * hdr[c * 4 + 1];
      float p75 = min_value + inc * hdr[c * 4 + 2], p100 = min_value + inc * hdr[c * 4 + 3];
      const uint8_t* col = bytes.data() + static_cast<size_t>(c) * rows;  // column-major bytes
      for (int32_t r = 0; r < rows; ++r) {
        int v = col[r];
        float f;
        if (v <= 64) f = p0 + (p25 - p0) * v * (1 / 64.0f);
        else if (v <= 192) f = p25 + (p75 - p25) * (v - 64) * (1 / 128.0f);
        else f = p75 + (p100 - p75) * (v - 192) * (1 / 63.0f);
        m->data[static_cast<size_t>(r) * cols + c] = f;
      }
    }
  } else if (tag == "CM2") {
    std::vector<uint16_t> w(n);
    in.Read(w.data(), n * 2);
    const float inc = range * (1.0f / 65535.0f);
    for (size_t i = 0; i < n; ++i) m->data[i] = min_value + inc * w[i];
  } else {
    std::vector<uint8_t> b(n);
    in.Read(b.data(), n);
    const float inc = range * (1.0f / 255.0f);
    for (size_t i = 0; i < n; ++i) m->data[i] = min_value + inc * b[i];
  }
}

void ReadVectorOrMatrix(Input& in, bool binary, Matrix* m, bool* was_vector) {
  m->ext = nullptr;
  if (!binary) {
    ReadTextMatrix(in, m);
    if (was_vector) *was_vector = (m->rows <= 1);
    return;
  }
  std::string tag = in.ReadToken();
  if (tag == "FV" || tag == "DV") {
    int32_t n = ReadInt32(in, true);
    if (n < 0) throw KaldiError("negative vector size in " + in.Name());
    m->rows = 1;
    m->cols = n;
    m->data.resize(n);
    if (tag == "FV") {
      in.Read(m->data.data(), static_cast<size_t>(n) * 4);
    } else {
      std::vector<double> d(n);
      in.Read(d.data(), static_cast<size_t>(n) * 8);
      for (int32_t i = 0; i < n; ++i) m->data[i] = static_cast<float>(d[i]);
    }
    if (was_vector) *was_vector = true;
    return;
  }
  if (was_vector) *was_vector = false;
  if (tag == "FM" || tag == "DM") {
    int32_t rows = ReadInt32(in, true), cols = ReadInt32(in, true);
    if (rows < 0 || cols < 0) throw KaldiError("negative matrix size in " + in.Name());
    m->rows = rows;
    m->cols = cols;
    const size_t n = static_cast<size_t>(rows) * cols;
    m->ext = (tag == "FM" && m->placer) ? m->placer(rows, cols) : nullptr;
    if (m->ext) {
      m->data.clear();
      in.Read(m->ext, n * 4);
      return;
    }
    m->data.resize(n);
    if (tag == "FM") {
      in.Read(m->data.data(), n * 4);
    } else {
      std::vector<double> d(n);
      in.Read(d.data(), n * 8);
      for (size_t i = 0; i < n; ++i) m->data[i] = static_cast<float>(d[i]);
    }
    return;
  }
  if (tag == "CM" || tag == "CM2" || tag == "CM3") {
    ReadCompressed(in, tag, m);
    return;
  }
  throw KaldiError("unknown matrix/vector tag '" + tag + "' in " + in.Name());
}

void ReadVectorOrMatrixDouble(Input& in, bool binary, int32_t* rows, int32_t* cols, std::vector<double>* data,
                              bool* was_vector) {
  data->clear();
  if (!binary) {
    // text "[ ... ]": rows end at '\n' (or at ']'); parsed in double
    in.SkipSpace();
    if (in.Get() != '[') throw KaldiError("expected '[' in " + in.Name());
    int32_t r = 0, c = 0, cur = 0;
    std::string tok;
    while (true) {
      int ch = in.Get();
      if (ch < 0) throw KaldiError("unterminated text matrix in " + in.Name());
      if (std::isspace(ch) || ch == ']') {
        if (!tok.empty()) {
          char* e = nullptr;
          double v = std::strtod(tok.c_str(), &e);
          if (*e) throw KaldiError("bad number '" + tok + "' in " + in.Name());
          data->push_back(v);
          tok.clear();
          ++cur;
        }
        if ((ch == '\n' || ch == ']') && cur > 0) {
          if (r == 0) c = cur;
          if (cur != c) throw KaldiError("ragged text matrix in " + in.Name());
          ++r;
          cur = 0;
        }
        if (ch == ']') break;
      } else {
        tok.push_back(static_cast<char>(ch));
      }
    }
    *rows = r;
    *cols = c;
    if (was_vector) *was_vector = r <= 1;
    return;
  }
  const std::string tag = in.ReadToken();
  const bool vec = tag == "FV" || tag == "DV";
  if (!vec && tag != "FM" && tag != "DM") throw KaldiError("unknown matrix/vector tag '" + tag + "' in " + in.Name());
  *rows = vec ? 1 : ReadInt32(in, true);
  *cols = ReadInt32(in, true);
  if (*rows < 0 || *cols < 0) throw KaldiError("negative size in " + in.Name());
  const size_t n = static_cast<size_t>(*rows) * *cols;
  data->resize(n);
  if (tag[0] == 'D') {
    in.Read(data->data(), n * 8);
  } else {
    std::vector<float> f(n);
    in.Read(f.data(), n * 4);
    for (size_t i = 0; i < n; ++i) (*data)[i] = f[i];
  }
  if (was_vector) *was_vector = vec;
}

void ReadMatrix(Input& in, bool binary, Matrix* m) {
  bool was_vector = false;
  ReadVectorOrMatrix(in, binary, m, &was_vector);
  if (binary && was_vector) throw KaldiError("expected a matrix, got a vector in " + in.Name());
}

void ReadVector(Input& in, bool binary, std::vector<float>* v) {
  Matrix m;
  bool was_vector = false;
  ReadVectorOrMatrix(in, binary, &m, &was_vector);
  if (m.rows > 1) throw KaldiError("expected a vector, got a matrix in " + in.Name());
  v->swap(m.data);
}

void WriteToken(Output& out, const std::string& tok) {
  out.Write(tok);
  out.Write(" ", 1);
}

static void WriteInt32Binary(Output& out, int32_t v) {
  char c = 4;
  out.Write(&c, 1);
  out.Write(&v, 4);
}

static void AppendFloatText(std::string* s, float v) {
  char b[32];
  if (std::isinf(v)) std::snprintf(b, sizeof b, v > 0 ? "inf" : "-inf");
  else if (std::isnan(v)) std::snprintf(b, sizeof b, "nan");
  else std::snprintf(b, sizeof b, "%.9g", v);
  s->append(b);
}

void WriteVector(Output& out, bool binary, const float* v, int32_t n) {
  if (binary) {
    out.Write("FV ", 3);
    WriteInt32Binary(out, n);
    out.Write(v, static_cast<size_t>(n) * 4);
  } else {
    std::string s = " [ ";
    for (int32_t i = 0; i < n; ++i) {
      AppendFloatText(&s, v[i]);
      s.push_back(' ');
    }
    s += "]\n";
    out.Write(s);
  }
}

void WriteMatrix(Output& out, bool binary, const Matrix& m) {
  WriteMatrix(out, binary, m.Data(), m.rows, m.cols);
}

void WriteMatrix(Output& out, bool binary, const float* data, int32_t rows, int32_t cols) {
  if (binary) {
    out.Write("FM ", 3);
    WriteInt32Binary(out, rows);
    WriteInt32Binary(out, cols);
    out.Write(data, static_cast<size_t>(rows) * cols * 4);
  } else {
    if (rows == 0) {
      out.Write(" [ ]\n");
      return;
    }
    std::string s = " [";
    for (int32_t r = 0; r < rows; ++r) {
      s += "\n  ";
      for (int32_t c = 0; c < cols; ++c) {
        AppendFloatText(&s, data[static_cast<size_t>(r) * cols + c]);
        s.push_back(' ');
      }
      if (s.size() > (1u << 20)) {  // keep the text buffer bounded for long utterances
        out.Write(s);
        s.clear();
      }
    }
    s += "]\n";
    out.Write(s);
  }
}

// --------------------------------------------------------------- tables ----

static TableSpec ParseSpec(const std::string& spec, bool is_write) {
  size_t colon = spec.find(':');
  if (colon == std::string::npos) throw KaldiError("invalid table specifier (no ':'): " + spec);
  TableSpec ts;
  std::string opts = spec.substr(0, colon), rest = spec.substr(colon + 1);
  std::vector<std::string> order;
  size_t p = 0;
  while (p <= opts.size()) {
    size_t q = opts.find(',', p);
    if (q == std::string::npos) q = opts.size();
    std::string o = opts.substr(p, q - p);
    p = q + 1;
    if (o == "ark") { ts.is_ark = true; order.push_back(o); }
    else if (o == "scp") { ts.is_scp = true; order.push_back(o); }
    else if (o == "t") ts.text = true;
    else if (o == "b") ts.text = false;
    else if (o == "p") ts.permissive = true;
    else if (o == "s") ts.sorted = true;
    else if (o == "cs") ts.called_sorted = true;
    else if (o == "o") ts.once = true;
    else if (o == "f") ts.flush = true;
    else if (o == "ns" || o == "ncs" || o == "no" || o == "np" || o == "nf" || o == "bg") {}
    else throw KaldiError("invalid option '" + o + "' in table specifier " + spec);
  }
  if (!ts.is_ark && !ts.is_scp) throw KaldiError("table specifier needs ark or scp: " + spec);
  if (ts.is_ark && ts.is_scp) {
    if (!is_write) throw KaldiError("ark,scp is only valid for writing: " + spec);
    if (order[0] != "ark") throw KaldiError("write specifier must be ark,scp (not scp,ark): " + spec);
    size_t comma = rest.find(',');
    if (comma == std::string::npos) throw KaldiError("ark,scp needs two file names: " + spec);
    ts.ark_name = Trim(rest.substr(0, comma));
    ts.scp_name = Trim(rest.substr(comma + 1));
  } else if (ts.is_ark) {
    ts.ark_name = Trim(rest);
  } else {
    ts.scp_name = Trim(rest);
  }
  return ts;
}

TableSpec ParseRspecifier(const std::string& s) { return ParseSpec(s, false); }
TableSpec ParseWspecifier(const std::string& s) { return ParseSpec(s, true); }

SequentialMatrixReader::SequentialMatrixReader(const std::string& rspecifier, bool allow_vectors,
                                               std::function<bool(const std::string&)> keep)
    : keep_(std::move(keep)), spec_(ParseRspecifier(rspecifier)), allow_vectors_(allow_vectors) {
  if (spec_.is_ark) ark_.Open(spec_.ark_name);
  else scp_.Open(spec_.scp_name);
  Advance();
}

void SequentialMatrixReader::Advance() {
  have_ = false;
  if (spec_.is_ark) {
    while (true) {
      ark_.SkipSpace();
      if (ark_.Eof()) { done_ = true; return; }
      key_ = ark_.ReadToken();
      bool binary = ReadBinaryHeader(ark_);
      const bool keep = !keep_ || keep_(key_);
      auto placer = value_.placer;
      if (!keep) value_.placer = nullptr;  // a skipped entry must not land in the caller's buffer
      if (allow_vectors_) ReadVectorOrMatrix(ark_, binary, &value_, nullptr);
      else ReadMatrix(ark_, binary, &value_);
      value_.placer = placer;
      if (!keep) continue;
      have_ = true;
      return;
    }
  }
  while (true) {
    if (scp_.Eof()) { done_ = true; return; }
    std::string line = Trim(scp_.ReadLine());
    if (line.empty()) continue;
    size_t sp = line.find_first_of(" \t");
    if (sp == std::string::npos) throw KaldiError("bad scp line: " + line);
    key_ = line.substr(0, sp);
    if (keep_ && !keep_(key_)) continue;
    std::string rx = Trim(line.substr(sp + 1));
    if (!rx.empty() && rx.back() == ']') throw KaldiError("scp range specifiers are not supported: " + line);
    Input in(rx);
    bool binary = ReadBinaryHeader(in);
    if (allow_vectors_) ReadVectorOrMatrix(in, binary, &value_, nullptr);
    else ReadMatrix(in, binary, &value_);
    have_ = true;
    return;
  }
}

std::vector<std::pair<std::string, std::string>> SequentialMatrixReader::ReadScpList(const std::string& scp_rxfilename) {
  std::vector<std::pair<std::string, std::string>> out;
  Input in(scp_rxfilename);
  while (!in.Eof()) {
    std::string line = Trim(in.ReadLine());
    if (line.empty()) continue;
    size_t sp = line.find_first_of(" \t");
    if (sp == std::string::npos) throw KaldiError("bad scp line: " + line);
    out.emplace_back(line.substr(0, sp), Trim(line.substr(sp + 1)));
  }
  in.Close();
  return out;
}

bool SequentialMatrixReader::Done() { return done_ && !have_; }
void SequentialMatrixReader::Next() { Advance(); }
int SequentialMatrixReader::Close() {
  int s = spec_.is_ark ? ark_.Close() : scp_.Close();
  done_ = true;
  have_ = false;
  return s;
}

VectorWriter::VectorWriter(const std::string& wspecifier) : spec_(ParseWspecifier(wspecifier)) {
  if (spec_.is_ark) ark_.Open(spec_.ark_name);
  if (spec_.is_scp) scp_.Open(spec_.scp_name);
  if (spec_.is_scp && !spec_.is_ark)
    throw KaldiError("writing an scp without an ark is not supported: " + wspecifier);
}

void VectorWriter::Write(const std::string& key, const float* v, int32_t n) {
  ark_.Write(key);
  ark_.Write(" ", 1);
  int64_t off = ark_.Tell();
  if (!spec_.text) ark_.Write("\0B", 2);
  WriteVector(ark_, !spec_.text, v, n);
  if (spec_.is_scp) {
    std::string line = key + " " + spec_.ark_name + ":" + std::to_string(off) + "\n";
    scp_.Write(line);
  }
  if (spec_.flush) {
    ark_.Flush();
    if (spec_.is_scp) scp_.Flush();
  }
}

MatrixWriter::MatrixWriter(const std::string& wspecifier) : spec_(ParseWspecifier(wspecifier)) {
  if (spec_.is_ark) ark_.Open(spec_.ark_name);
  if (spec_.is_scp) scp_.Open(spec_.scp_name);
  if (spec_.is_scp && !spec_.is_ark)
    throw KaldiError("writing an scp without an ark is not supported: " + wspecifier);
}

void MatrixWriter::Write(const std::string& key, const float* data, int32_t rows, int32_t cols) {
  ark_.Write(key);
  ark_.Write(" ", 1);
  int64_t off = ark_.Tell();
  if (!spec_.text) ark_.Write("\0B", 2);
  WriteMatrix(ark_, !spec_.text, data, rows, cols);
  if (spec_.is_scp) {
    std::string line = key + " " + spec_.ark_name + ":" + std::to_string(off) + "\n";
    scp_.Write(line);
  }
  if (spec_.flush) {
    ark_.Flush();
    if (spec_.is_scp) scp_.Flush();
  }
}

void MatrixWriter::Close() {
  if (ark_.IsOpen() && ark_.Close() != 0) throw KaldiError("error closing " + spec_.ark_name);
  if (scp_.IsOpen() && scp_.Close() != 0) throw KaldiError("error closing " + spec_.scp_name);
}

void VectorWriter::Close() {
  if (ark_.IsOpen() && ark_.Close() != 0) throw KaldiError("error closing " + spec_.ark_name);
  if (scp_.IsOpen() && scp_.Close() != 0) throw KaldiError("error closing " + spec_.scp_name);
}

// --------------------------------------------------------- ParseOptions ----

std::string ParseOptions::Normalize(const std::string& name) {
  std::string o;
  for (char c : name) o.push_back(c == '_' ? '-' : static_cast<char>(std::tolower(static_cast<unsigned char>(c))));
  return o;
}

void ParseOptions::Register(const std::string& n, bool* v, const std::string& d) { opts_[Normalize(n)] = {kBool, v, d}; }
void ParseOptions::Register(const std::string& n, int32_t* v, const std::string& d) { opts_[Normalize(n)] = {kInt32, v, d}; }
void ParseOptions::Register(const std::string& n, int64_t* v, const std::string& d) { opts_[Normalize(n)] = {kInt64, v, d}; }
void ParseOptions::Register(const std::string& n, float* v, const std::string& d) { opts_[Normalize(n)] = {kFloat, v, d}; }
void ParseOptions::Register(const std::string& n, std::string* v, const std::string& d) { opts_[Normalize(n)] = {kString, v, d}; }

static bool ParseBool(const std::string& v, const std::string& key) {
  std::string s;
  for (char c : v) s.push_back(static_cast<char>(std::tolower(static_cast<unsigned char>(c))));
  if (s == "true" || s == "t" || s == "1" || s.empty()) return true;
  if (s == "false" || s == "f" || s == "0") return false;
  throw KaldiError("invalid boolean value '" + v + "' for option --" + key);
}

void ParseOptions::SetOption(const std::string& key_in, const std::string& value, bool has_value) {
  std::string key = Normalize(key_in);
  if (key == "help") {
    PrintUsage(true);
    std::exit(0);
  }
  if (key == "verbose") { verbose = std::atoi(value.c_str()); return; }
  if (key == "print-args") { print_args_ = ParseBool(value, key); return; }
  if (key == "config") { ReadConfigFile(value); return; }
  auto it = opts_.find(key);
  if (it == opts_.end()) throw KaldiError("invalid option --" + key_in);
  Opt& o = it->second;
  if (o.kind != kBool && !has_value) throw KaldiError("option --" + key + " needs a value");
  char* e = nullptr;
  switch (o.kind) {
    case kBool: *static_cast<bool*>(o.ptr) = ParseBool(value, key); break;
    case kInt32: {
      long v = std::strtol(value.c_str(), &e, 10);
      if (value.empty() || *e) throw KaldiError("invalid integer '" + value + "' for option --" + key);
      *static_cast<int32_t*>(o.ptr) = static_cast<int32_t>(v);
      break;
    }
    case kInt64: {
      long long v = std::strtoll(value.c_str(), &e, 10);
      if (value.empty() || *e) throw KaldiError("invalid integer '" + value + "' for option --" + key);
      *static_cast<int64_t*>(o.ptr) = v;
      break;
    }
    case kFloat: {
      float v = std::strtof(value.c_str(), &e);
      if (value.empty() || *e) throw KaldiError("invalid float '" + value + "' for option --" + key);
      *static_cast<float*>(o.ptr) = v;
      break;
    }
    case kString: *static_cast<std::string*>(o.ptr) = value; break;
  }
}

void ParseOptions::ReadConfigFile(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw KaldiError("cannot open config file " + path);
  std::string line;
  while (std::getline(f, line)) {
    size_t h = line.find('#');
    if (h != std::string::npos) line = line.substr(0, h);
    line = Trim(line);
    if (line.empty()) continue;
    if (line.compare(0, 2, "--") != 0) throw KaldiError("bad line in config file " + path + ": " + line);
    size_t eq = line.find('=');
    if (eq == std::string::npos) SetOption(line.substr(2), "", false);
    else SetOption(line.substr(2, eq - 2), line.substr(eq + 1), true);
  }
}

int ParseOptions::Read(int argc, const char* const* argv) {
  if (argc > 0) SetProgramName(argv[0]);
  int i = 1;
  for (; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--") { ++i; break; }
    if (a.compare(0, 2, "--") != 0) break;
    size_t eq = a.find('=');
    if (eq == std::string::npos) SetOption(a.substr(2), "", false);
    else SetOption(a.substr(2, eq - 2), a.substr(eq + 1), true);
  }
  for (; i < argc; ++i) args_.push_back(argv[i]);
  if (print_args_) {
    std::string s;
    for (int k = 0; k < argc; ++k) {
      std::string a = argv[k];
      bool quote = a.find_first_of(" \t|'\"") != std::string::npos;
      if (k) s.push_back(' ');
      s += quote ? ("'" + a + "'") : a;
    }
    std::fprintf(stderr, "%s\n", s.c_str());
  }
  return NumArgs();
}

const std::string& ParseOptions::GetArg(int i) const {
  if (i < 1 || i > NumArgs()) throw KaldiError("GetArg: index out of range");
  return args_[i - 1];
}

void ParseOptions::PrintUsage(bool to_stderr) const {
  FILE* f = to_stderr ? stderr : stdout;
  std::fprintf(f, "\n%s\n\nOptions:\n", usage_.c_str());
  for (const auto& kv : opts_) {
    const Opt& o = kv.second;
    std::string def;
    switch (o.kind) {
      case kBool: def = *static_cast<bool*>(o.ptr) ? "true" : "false"; break;
      case kInt32: def = std::to_string(*static_cast<int32_t*>(o.ptr)); break;
      case kInt64: def = std::to_string(*static_cast<int64_t*>(o.ptr)); break;
      case kFloat: def = std::to_string(*static_cast<float*>(o.ptr)); break;
      case kString: def = "\"" + *static_cast<std::string*>(o.ptr) + "\""; break;
    }
    std::fprintf(f, "  --%-24s : %s (default = %s)\n", kv.first.c_str(), o.doc.c_str(), def.c_str());
  }
  std::fprintf(f, "\nStandard options:\n  --config, --help, --print-args, --verbose\n\n");
}

}  // namespace cvx
