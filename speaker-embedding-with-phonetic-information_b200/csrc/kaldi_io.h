// Kaldi I/O surface of the extraction programs: rxfilename/wxfilename (file,
// "-", pipes), binary/text tokens, float matrices/vectors (incl. compressed
// CM/CM2/CM3), and sequential table readers / writers for
// `ark:` / `scp:` / `ark,scp:` / `ark,t:` specifiers.
//
// Contract taken from the reference's call sites (Kaldi itself is not shipped,
// SURVEY.md section 0):
//   features   egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:79   "ark:apply-cmvn-sliding ... |"
//   output     extract_xvectors_new.sh:93                               "ark,scp:ARK,SCP"
//   model      extract_xvectors_new.sh:59                               "nnet3-copy ... - |"
#ifndef CVX_KALDI_IO_H_
#define CVX_KALDI_IO_H_

#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace cvx {

struct KaldiError : public std::runtime_error {
  using std::runtime_error::runtime_error;
};

// Kaldi-style logging on stderr ("LOG (prog:func():file:line) msg") so that
// utils/summarize_warnings.pl style tooling keeps working.
void SetProgramName(const std::string& name);
void LogMessage(const char* severity, const char* func, const char* file, int line,
                const std::string& msg);
#define CVX_LOG(msg_expr)                                                        \
  do {                                                                           \
    std::ostringstream cvx_os__;                                                 \
    cvx_os__ << msg_expr;                                                        \
    ::cvx::LogMessage("LOG", __func__, __FILE__, __LINE__, cvx_os__.str());      \
  } while (0)
#define CVX_WARN(msg_expr)                                                       \
  do {                                                                           \
    std::ostringstream cvx_os__;                                                 \
    cvx_os__ << msg_expr;                                                        \
    ::cvx::LogMessage("WARNING", __func__, __FILE__, __LINE__, cvx_os__.str());  \
  } while (0)

// ---------------------------------------------------------------- streams --
// Buffered byte source over a file, stdin or a popen()'d command.
class Input {
 public:
  Input() = default;
  explicit Input(const std::string& rxfilename) { Open(rxfilename); }
  ~Input() { Close(); }
  Input(const Input&) = delete;
  Input& operator=(const Input&) = delete;

  // rxfilename: "path", "path:offset", "-" or "cmd |".
  void Open(const std::string& rxfilename);
  // Returns the exit status of a piped command (0 for files).
  int Close();
  bool IsOpen() const { return fp_ != nullptr; }

  int Peek();  // -1 at EOF
  int Get();
  void Read(void* dst, size_t n);  // throws on short read
  bool Eof() { return Peek() < 0; }
  void SkipSpace();
  // Kaldi ReadToken: skips leading space, reads to next space, consumes one space.
  std::string ReadToken();
  std::string PeekToken();  // text lookahead of one token (bounded by buffer)
  std::string ReadLine();   // up to and excluding '\n'
  const std::string& Name() const { return name_; }

 private:
  bool Fill();
  FILE* fp_ = nullptr;
  bool is_pipe_ = false, is_stdin_ = false;
  std::string name_;
  std::vector<unsigned char> buf_;
  size_t pos_ = 0, end_ = 0;
};

class Output {
 public:
  Output() = default;
  ~Output() { Close(); }
  Output(const Output&) = delete;
  Output& operator=(const Output&) = delete;
  // wxfilename: "path", "-" or "| cmd".
  void Open(const std::string& wxfilename);
  int Close();
  bool IsOpen() const { return fp_ != nullptr; }
  void Write(const void* src, size_t n);
  void Write(const std::string& s) { Write(s.data(), s.size()); }
  int64_t Tell() const { return bytes_; }
  void Flush();
  const std::string& Name() const { return name_; }

 private:
  FILE* fp_ = nullptr;
  bool is_pipe_ = false, is_stdout_ = false;
  std::string name_;
  int64_t bytes_ = 0;
};

// ------------------------------------------------------------ basic types --
bool ReadBinaryHeader(Input& in);  // consumes "\0B" if present, returns binary?
void ExpectToken(Input& in, const std::string& tok);
int32_t ReadInt32(Input& in, bool binary);
float ReadFloat(Input& in, bool binary);  // accepts 4- or 8-byte binary reals
bool ReadBool(Input& in, bool binary);

struct Matrix {
  int32_t rows = 0, cols = 0;
  std::vector<float> data;  // row-major (unless `ext` is set)
  // Optional placement: called with (rows, cols) of an uncompressed binary float matrix before
  // its body is read; a non-null result receives the data directly (e.g. the tail of a pinned
  // staging buffer) and `ext` points at it.
  std::function<float*(int32_t, int32_t)> placer;
  float* ext = nullptr;
  const float* Data() const { return ext ? ext : data.data(); }
  float* Row(int32_t r) { return (ext ? ext : data.data()) + static_cast<size_t>(r) * cols; }
  const float* Row(int32_t r) const { return Data() + static_cast<size_t>(r) * cols; }
};

// After an optional binary header: FM/DM/CM/CM2/CM3 (binary) or "[ ... ]" (text).
void ReadMatrix(Input& in, bool binary, Matrix* m);
// FV/DV (binary) or "[ ... ]" (text).
void ReadVector(Input& in, bool binary, std::vector<float>* v);
// Reads FV/DV/FM/DM/[..] without knowing which; vectors come back as 1 x n.
void ReadVectorOrMatrix(Input& in, bool binary, Matrix* m, bool* was_vector);

// Double-preserving read of a vector (1 x n) or matrix: DV/DM stay doubles, FV/FM and text are
// widened (text is parsed with strtod).  What Kaldi's Vector<double>/Matrix<double>::Read does for
// objects stored in double precision, e.g. <Plda>.
void ReadVectorOrMatrixDouble(Input& in, bool binary, int32_t* rows, int32_t* cols, std::vector<double>* data,
                              bool* was_vector);

void WriteToken(Output& out, const std::string& tok);
void WriteVector(Output& out, bool binary, const float* v, int32_t n);
void WriteMatrix(Output& out, bool binary, const Matrix& m);
void WriteMatrix(Output& out, bool binary, const float* data, int32_t rows, int32_t cols);

// ----------------------------------------------------------------- tables --
struct TableSpec {
  bool is_ark = false, is_scp = false;  // wspecifier may have both
  bool text = false;                    // ",t"
  bool permissive = false, sorted = false, called_sorted = false, once = false, flush = false;
  std::string ark_name, scp_name;       // for "ark,scp:A,B"; else the single name in the matching one
};
TableSpec ParseRspecifier(const std::string& rspecifier);
TableSpec ParseWspecifier(const std::string& wspecifier);

// SequentialBaseFloatMatrixReader equivalent.
class SequentialMatrixReader {
 public:
  // allow_vectors: entries may be Kaldi vectors (returned as 1 x n), e.g. a vad.scp table
  explicit SequentialMatrixReader(const std::string& rspecifier, bool allow_vectors = false,
                                  std::function<bool(const std::string&)> keep = nullptr);
  bool Done();            // true when no more entries
  const std::string& Key() const { return key_; }
  Matrix& Value() { return value_; }
  void Next();
  // Exit status problems of the feeding pipe are reported at Close().
  int Close();
  // (`keep` of the constructor: entries whose key it rejects are skipped -- scp lines without opening
  // their file, archive entries after reading them.)
  // scp only: (key, rxfilename) of every line, without reading any entry
  static std::vector<std::pair<std::string, std::string>> ReadScpList(const std::string& scp_rxfilename);
  const TableSpec& Spec() const { return spec_; }

 private:
  void Advance();
  std::function<bool(const std::string&)> keep_;
  TableSpec spec_;
  Input ark_;                 // for ark
  Input scp_;                 // for scp: the list itself
  std::string key_;
  Matrix value_;
  bool have_ = false, done_ = false, allow_vectors_ = false;
};

// BaseFloatVectorWriter equivalent for ark / ark,t / ark,scp / scp-less forms.
class VectorWriter {
 public:
  explicit VectorWriter(const std::string& wspecifier);
  void Write(const std::string& key, const float* v, int32_t n);
  void Close();

 private:
  TableSpec spec_;
  Output ark_;
  Output scp_;
};

// Table of float matrices (what nnet3-compute writes): ark, ark,t and ark,scp.
class MatrixWriter {
 public:
  explicit MatrixWriter(const std::string& wspecifier);
  void Write(const std::string& key, const float* data, int32_t rows, int32_t cols);
  void Close();

 private:
  TableSpec spec_;
  Output ark_;
  Output scp_;
};

// ---------------------------------------------------------- option parser --
// Kaldi ParseOptions subset: --name=value, bare --flag for bools, --config=file,
// --help, --verbose, --print-args; positional args after options.
class ParseOptions {
 public:
  explicit ParseOptions(const std::string& usage) : usage_(usage) {}
  void Register(const std::string& name, bool* v, const std::string& doc);
  void Register(const std::string& name, int32_t* v, const std::string& doc);
  void Register(const std::string& name, int64_t* v, const std::string& doc);
  void Register(const std::string& name, float* v, const std::string& doc);
  void Register(const std::string& name, std::string* v, const std::string& doc);
  // Returns number of positional args; exits(0) on --help; throws KaldiError on bad options.
  int Read(int argc, const char* const* argv);
  int NumArgs() const { return static_cast<int>(args_.size()); }
  const std::string& GetArg(int i) const;  // 1-based, like Kaldi
  void PrintUsage(bool to_stderr = true) const;
  int verbose = 0;

 private:
  enum Kind { kBool, kInt32, kInt64, kFloat, kString };
  struct Opt {
    Kind kind;
    void* ptr;
    std::string doc;
  };
  void SetOption(const std::string& key, const std::string& value, bool has_value);
  void ReadConfigFile(const std::string& path);
  static std::string Normalize(const std::string& name);
  std::string usage_;
  std::map<std::string, Opt> opts_;
  std::vector<std::string> args_;
  bool print_args_ = true;
};

}  // namespace cvx

#include <sstream>
#endif  // CVX_KALDI_IO_H_
