// nnet3 raw model -> fused-layer Plan.
//
// File layout restated from upstream Kaldi src/nnet3/nnet-nnet.cc (Nnet::Read),
// nnet-simple-component.cc, nnet-normalize-component.cc,
// nnet-general-component.cc (Kaldi is not shipped by
// the reference).  What the reference itself pins are the node/component lines
// that end up in final.raw:
//   egs/sre/v5/steps/libs/nnet3/xconfig/basic_layers.py:758-834  (affine -> relu -> batchnorm)
//   egs/sre/v5/steps/libs/nnet3/xconfig/stats_layer.py:96-132    (extraction/pooling, Round(.., 1))
// Components are read token by token and unknown tokens are skipped generically,
// so version-dependent trailing fields (natural-gradient state etc.) do not matter.
#include "nnet3_model.h"

#include <cctype>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <sstream>

#include "kaldi_io.h"

namespace cvx {
namespace {

// ------------------------------------------------------------ descriptors --
struct Desc {
  enum Kind { kNode, kOffset, kAppend, kRound } kind = kNode;
  std::string node;
  int value = 0;            // Offset: t offset; Round: modulus
  std::vector<Desc> args;   // Append: parts; Offset/Round: one child
};

class DescParser {
 public:
  explicit DescParser(const std::string& s) : src_(s) {
    std::string cur;
    for (char c : s) {
      if (c == '(' || c == ')' || c == ',') {
        if (!cur.empty()) toks_.push_back(cur), cur.clear();
        toks_.push_back(std::string(1, c));
      } else if (std::isspace(static_cast<unsigned char>(c))) {
        if (!cur.empty()) toks_.push_back(cur), cur.clear();
      } else {
        cur.push_back(c);
      }
    }
    if (!cur.empty()) toks_.push_back(cur);
  }
  Desc Parse() {
    Desc d = ParseOne();
    if (pos_ != toks_.size()) throw FormatError("trailing tokens in descriptor '" + src_ + "'");
    return d;
  }

 private:
  static bool IsInt(const std::string& t) {
    size_t i = (t[0] == '-' || t[0] == '+') ? 1 : 0;
    if (i >= t.size()) return false;
    for (; i < t.size(); ++i)
      if (!std::isdigit(static_cast<unsigned char>(t[i]))) return false;
    return true;
  }
  const std::string& Tok() {
    if (pos_ >= toks_.size()) throw FormatError("unexpected end of descriptor '" + src_ + "'");
    return toks_[pos_];
  }
  void Expect(const std::string& t) {
    if (Tok() != t) throw FormatError("expected '" + t + "' in descriptor '" + src_ + "'");
    ++pos_;
  }
  Desc ParseOne() {
    std::string t = Tok();
    ++pos_;
    if (pos_ < toks_.size() && toks_[pos_] == "(") {
      ++pos_;
      Desc d;
      if (t == "Append") {
        d.kind = Desc::kAppend;
        while (true) {
          d.args.push_back(ParseOne());
          if (Tok() == ",") { ++pos_; continue; }
          Expect(")");
          break;
        }
        if (d.args.size() == 1) return d.args[0];
        return d;
      }
      if (t == "Offset" || t == "Round") {
        d.kind = (t == "Offset") ? Desc::kOffset : Desc::kRound;
        d.args.push_back(ParseOne());
        Expect(",");
        if (!IsInt(Tok())) throw FormatError("expected integer in descriptor '" + src_ + "'");
        d.value = std::atoi(Tok().c_str());
        ++pos_;
        if (t == "Offset" && Tok() == ",") {  // Offset(d, t, x)
          ++pos_;
          if (!IsInt(Tok()) || std::atoi(Tok().c_str()) != 0)
            throw UnsupportedError("Offset with an x offset is not on the extraction path: " + src_);
          ++pos_;
        }
        Expect(")");
        return d;
      }
      throw UnsupportedError("descriptor function '" + t + "' is not produced by the reference's recipes: " + src_);
    }
    Desc d;
    d.kind = Desc::kNode;
    d.node = t;
    return d;
  }
  std::string src_;
  std::vector<std::string> toks_;
  size_t pos_ = 0;
};

// ---------------------------------------------------------- model in file --
struct Value {
  enum Kind { kScalar4, kDouble, kBool, kMatrix, kText } kind = kText;
  unsigned char raw[4] = {0, 0, 0, 0};
  double d = 0;
  bool b = false;
  Matrix m;
  bool is_vector = false;
  std::string text;
};

struct Component {
  std::string type;
  std::map<std::string, Value> fields;

  const Value* Find(const std::string& k) const {
    auto it = fields.find(k);
    return it == fields.end() ? nullptr : &it->second;
  }
  const Value& Get(const std::string& k, const std::string& cname) const {
    const Value* v = Find(k);
    if (!v) throw FormatError("component " + cname + " (" + type + ") has no <" + k + ">");
    return *v;
  }
  int32_t Int(const std::string& k, const std::string& cname) const {
    const Value& v = Get(k, cname);
    if (v.kind == Value::kScalar4) { int32_t i; std::memcpy(&i, v.raw, 4); return i; }
    if (v.kind == Value::kText) return static_cast<int32_t>(std::strtol(v.text.c_str(), nullptr, 10));
    if (v.kind == Value::kDouble) return static_cast<int32_t>(v.d);
    throw FormatError("<" + k + "> of " + cname + " is not an integer");
  }
  double Real(const std::string& k, const std::string& cname) const {
    const Value& v = Get(k, cname);
    if (v.kind == Value::kScalar4) { float f; std::memcpy(&f, v.raw, 4); return f; }
    if (v.kind == Value::kDouble) return v.d;
    if (v.kind == Value::kText) return std::strtod(v.text.c_str(), nullptr);
    throw FormatError("<" + k + "> of " + cname + " is not a number");
  }
  bool Bool(const std::string& k, const std::string& cname) const {
    const Value& v = Get(k, cname);
    if (v.kind == Value::kBool) return v.b;
    if (v.kind == Value::kText) return v.text == "T" || v.text == "true";
    throw FormatError("<" + k + "> of " + cname + " is not a bool");
  }
  const Matrix& Mat(const std::string& k, const std::string& cname) const {
    const Value& v = Get(k, cname);
    if (v.kind != Value::kMatrix) throw FormatError("<" + k + "> of " + cname + " is not a matrix/vector");
    return v.m;
  }
};

struct Node {
  enum Kind { kInput, kComponent, kOutput } kind = kInput;
  std::string name;
  int32_t dim = 0;
  std::string component;
  Desc desc;
};

struct Nnet {
  std::vector<Node> nodes;
  std::map<std::string, size_t> node_index;
  std::map<std::string, Component> components;
};

std::map<std::string, std::string> ParseKeyValues(const std::string& rest, const std::string& line) {
  // values may contain spaces (descriptors): a new key starts at " name="
  std::map<std::string, std::string> kv;
  size_t i = 0, n = rest.size();
  while (i < n) {
    while (i < n && std::isspace(static_cast<unsigned char>(rest[i]))) ++i;
    if (i >= n) break;
    size_t eq = rest.find('=', i);
    if (eq == std::string::npos) throw FormatError("bad config line: " + line);
    std::string key = rest.substr(i, eq - i);
    // value extends to the next " key=" where key is [A-Za-z][\w-]*
    size_t j = eq + 1, end = n;
    for (size_t p = j; p < n; ++p) {
      if (!std::isspace(static_cast<unsigned char>(rest[p]))) continue;
      size_t q = p;
      while (q < n && std::isspace(static_cast<unsigned char>(rest[q]))) ++q;
      size_t r = q;
      if (r < n && std::isalpha(static_cast<unsigned char>(rest[r]))) {
        while (r < n && (std::isalnum(static_cast<unsigned char>(rest[r])) || rest[r] == '-' || rest[r] == '_')) ++r;
        if (r < n && rest[r] == '=') { end = p; break; }
      }
      p = q > p ? q - 1 : p;
    }
    std::string val = rest.substr(j, end - j);
    while (!val.empty() && std::isspace(static_cast<unsigned char>(val.back()))) val.pop_back();
    kv[key] = val;
    i = end;
  }
  return kv;
}

void ReadGenericValue(Input& in, bool binary, Value* v) {
  if (binary) {
    int c = in.Peek();
    if (c == 4) {
      in.Get();
      in.Read(v->raw, 4);
      v->kind = Value::kScalar4;
      return;
    }
    if (c == 8) {
      in.Get();
      in.Read(&v->d, 8);
      v->kind = Value::kDouble;
      return;
    }
    if (c == 'T') {
      in.Get();
      v->kind = Value::kBool;
      v->b = true;
      return;
    }
    if (c == 'F' || c == 'D' || c == 'C') {
      // 'F' alone is a bool; "FV "/"FM "/"DV "/"DM "/"CM*" start a vector/matrix.
      std::string t = in.PeekToken();
      if (t == "FV" || t == "FM" || t == "DV" || t == "DM" || t == "CM" || t == "CM2" || t == "CM3") {
        v->kind = Value::kMatrix;
        ReadVectorOrMatrix(in, true, &v->m, &v->is_vector);
        return;
      }
      if (c == 'F') {
        in.Get();
        v->kind = Value::kBool;
        v->b = false;
        return;
      }
    }
    throw FormatError("cannot parse a binary value in " + in.Name() + " (next byte " + std::to_string(c) + ")");
  }
  in.SkipSpace();
  if (in.Peek() == '[') {
    v->kind = Value::kMatrix;
    ReadVectorOrMatrix(in, false, &v->m, &v->is_vector);
    return;
  }
  std::string t = in.ReadToken();
  if (t == "T" || t == "F") {
    v->kind = Value::kBool;
    v->b = (t == "T");
    return;
  }
  v->kind = Value::kText;
  v->text = t;
}

Nnet ReadNnet(const std::string& rxfilename) {
  Input in(rxfilename);
  bool binary = ReadBinaryHeader(in);
  ExpectToken(in, "<Nnet3>");
  Nnet net;
  // config section: text lines, terminated by an empty line
  std::string line = in.ReadLine();  // rest of the "<Nnet3>" line
  bool first = true;
  while (true) {
    if (in.Eof()) throw FormatError("unexpected end of file in the config section of " + rxfilename);
    line = in.ReadLine();
    std::string t = line;
    while (!t.empty() && std::isspace(static_cast<unsigned char>(t.back()))) t.pop_back();
    size_t a = 0;
    while (a < t.size() && std::isspace(static_cast<unsigned char>(t[a]))) ++a;
    t = t.substr(a);
    if (t.empty()) {
      if (first) { first = false; continue; }  // tolerate one leading blank line
      break;
    }
    first = false;
    if (t[0] == '#') continue;
    size_t sp = t.find_first_of(" \t");
    std::string head = t.substr(0, sp), rest = sp == std::string::npos ? "" : t.substr(sp + 1);
    auto kv = ParseKeyValues(rest, t);
    Node nd;
    if (!kv.count("name")) throw FormatError("config line without name=: " + t);
    nd.name = kv["name"];
    if (head == "input-node") {
      nd.kind = Node::kInput;
      nd.dim = std::atoi(kv["dim"].c_str());
      if (nd.dim <= 0) throw FormatError("input-node without a valid dim: " + t);
    } else if (head == "component-node") {
      nd.kind = Node::kComponent;
      nd.component = kv["component"];
      nd.desc = DescParser(kv["input"]).Parse();
    } else if (head == "output-node") {
      nd.kind = Node::kOutput;
      nd.desc = DescParser(kv["input"]).Parse();
    } else if (head == "dim-range-node") {
      nd.kind = Node::kComponent;
      nd.component = "\x01" "dim-range";  // rejected only if reachable
    } else {
      throw FormatError("unknown config line type '" + head + "' in " + rxfilename);
    }
    net.node_index[nd.name] = net.nodes.size();
    net.nodes.push_back(std::move(nd));
  }
  ExpectToken(in, "<NumComponents>");
  int32_t n = ReadInt32(in, binary);
  if (n < 0 || n > 100000) throw FormatError("implausible <NumComponents> in " + rxfilename);
  for (int32_t i = 0; i < n; ++i) {
    ExpectToken(in, "<ComponentName>");
    std::string cname = in.ReadToken();
    std::string open = in.ReadToken();
    if (open.size() < 3 || open.front() != '<' || open.back() != '>')
      throw FormatError("expected a component type token after " + cname + ", got '" + open + "'");
    Component comp;
    comp.type = open.substr(1, open.size() - 2);
    const std::string close = "</" + comp.type + ">";
    while (true) {
      std::string tok = in.ReadToken();
      if (tok == close) break;
      if (tok.size() < 3 || tok.front() != '<' || tok.back() != '>')
        throw FormatError("component " + cname + ": expected a <Token>, got '" + tok + "'");
      Value v;
      ReadGenericValue(in, binary, &v);
      comp.fields[tok.substr(1, tok.size() - 2)] = std::move(v);
    }
    net.components[cname] = std::move(comp);
  }
  ExpectToken(in, "</Nnet3>");
  int status = in.Close();
  if (status != 0) throw KaldiError("model input pipe exited with status " + std::to_string(status));
  return net;
}

// ------------------------------------------------------------- lowering ----
bool IsAffine(const std::string& t) {
  return t == "NaturalGradientAffineComponent" || t == "AffineComponent" || t == "FixedAffineComponent" ||
         t == "LinearComponent";
}
bool IsIdentity(const std::string& t) {
  return t == "NoOpComponent" || t == "DropoutComponent" || t == "GeneralDropoutComponent";
}

// Symbolic value of a node while walking back from the output.
struct Sym {
  enum Kind { kInput, kFused, kExtraction, kPooled } kind = kInput;
  int32_t input = -1;            // kInput
  std::string affine_node;       // kFused: the component-node holding the affine
  std::string pre_bn;            // BatchNorm directly after the affine: folded into W, b
  bool relu = false;
  std::string post_bn;           // BatchNorm after ReLU (or after a plain affine)
  int32_t extraction_src = -1;   // kExtraction / kPooled: frame layer index
  std::string Key() const { return affine_node + "|" + pre_bn + "|" + (relu ? "r" : "") + "|" + post_bn; }
};

class Lowerer {
 public:
  Lowerer(const Nnet& net, Plan* plan, bool frame_output) : net_(net), plan_(plan), frame_output_(frame_output) {}

  void Run() {
    auto it = net_.node_index.find("output");
    if (it == net_.node_index.end()) throw FormatError("model has no output-node named 'output'");
    const Node& out = net_.nodes[it->second];
    if (out.kind != Node::kOutput) throw FormatError("'output' is not an output-node");
    for (const Node& n : net_.nodes)
      if (n.kind == Node::kInput) inputs_.push_back(n.name);
    if (frame_output_) {
      // nnet3-compute form: the output is a frame-level layer, optionally behind a LogSoftmax
      const Desc* d = &out.desc;
      if (d->kind == Desc::kNode) {
        const Node& n = NodeByName(d->node);
        if (n.kind == Node::kComponent && Comp(n).type == "LogSoftmaxComponent") {
          plan_->frame_log_softmax = true;
          d = &n.desc;
        }
      }
      Ref r = Flatten1(*d, "output");
      if (r.segment) throw UnsupportedError("the requested output is segment-level; use the extraction entry points");
      if (r.kind != Ref::kFrameLayer) throw UnsupportedError("the output must be a layer of the network, not an input");
      if (r.index != static_cast<int32_t>(plan_->frame.size()) - 1)
        throw UnsupportedError("internal: output layer is not the last lowered layer");
      plan_->frame_out = r.index;
    } else {
      Ref r = Flatten1(out.desc, "output");
      if (!r.segment)
        throw UnsupportedError("the requested output is frame-level; this path produces one vector per chunk");
      if (r.kind != Ref::kSegLayer && r.kind != Ref::kPool) throw UnsupportedError("unsupported output expression");
      if (plan_->pool_src < 0) throw UnsupportedError("no StatisticsPooling on the path to the output");
    }
    // keep only the inputs that are reachable, in file order
    std::vector<std::string> names;
    std::vector<int32_t> dims, remap(inputs_.size(), -1);
    for (size_t i = 0; i < inputs_.size(); ++i)
      if (used_inputs_.count(static_cast<int32_t>(i))) {
        remap[i] = static_cast<int32_t>(names.size());
        names.push_back(inputs_[i]);
        dims.push_back(net_.nodes[net_.node_index.at(inputs_[i])].dim);
      }
    for (Layer& l : plan_->frame)
      for (Tap& t : l.taps)
        if (t.src < 0) t.src = -1 - remap[-1 - t.src];
    plan_->input_names = names;
    plan_->input_dims = dims;
  }

 private:
  struct Ref {
    enum Kind { kInputStream, kFrameLayer, kPool, kSegLayer } kind = kInputStream;
    int32_t index = 0;  // stream / frame layer / segment layer
    bool segment = false;
  };

  const Node& NodeByName(const std::string& name) const {
    auto it = net_.node_index.find(name);
    if (it == net_.node_index.end()) throw FormatError("descriptor refers to unknown node '" + name + "'");
    return net_.nodes[it->second];
  }
  const Component& Comp(const Node& n) const {
    if (!n.component.empty() && n.component[0] == '\x01')
      throw UnsupportedError("dim-range-node '" + n.name + "' is not on the extraction path");
    auto it = net_.components.find(n.component);
    if (it == net_.components.end()) throw FormatError("node " + n.name + " uses unknown component " + n.component);
    return it->second;
  }

  Sym SymOf(const std::string& name) {
    auto m = sym_memo_.find(name);
    if (m != sym_memo_.end()) return m->second;
    if (in_progress_.count(name)) throw UnsupportedError("recurrent graph (cycle through " + name + ")");
    in_progress_.insert(name);
    const Node& n = NodeByName(name);
    Sym s;
    if (n.kind == Node::kInput) {
      s.kind = Sym::kInput;
      for (size_t i = 0; i < inputs_.size(); ++i)
        if (inputs_[i] == name) s.input = static_cast<int32_t>(i);
    } else if (n.kind == Node::kOutput) {
      throw UnsupportedError("descriptor reads from output-node " + name);
    } else {
      const Component& c = Comp(n);
      if (IsAffine(c.type)) {
        s.kind = Sym::kFused;
        s.affine_node = name;
      } else if (c.type == "RectifiedLinearComponent" || c.type == "BatchNormComponent" || IsIdentity(c.type)) {
        if (n.desc.kind != Desc::kNode)
          throw UnsupportedError(c.type + " node '" + name + "' must read a single node without offsets");
        s = SymOf(n.desc.node);
        if (IsIdentity(c.type)) {
          // test mode: identity
        } else if (s.kind != Sym::kFused) {
          throw UnsupportedError(c.type + " node '" + name + "' does not follow an affine component");
        } else if (c.type == "RectifiedLinearComponent") {
          if (s.relu || !s.post_bn.empty())
            throw UnsupportedError("unsupported component order before ReLU node '" + name + "'");
          s.relu = true;
        } else {  // BatchNorm: after the ReLU it becomes the epilogue's scale/offset,
                  // directly after the affine it is folded into W and b
          std::string& slot = s.relu ? s.post_bn : s.pre_bn;
          if (!slot.empty()) throw UnsupportedError("two BatchNorm components in a row at '" + name + "'");
          slot = n.component;
        }
      } else if (c.type == "StatisticsExtractionComponent") {
        if (c.Int("InputPeriod", n.component) != 1 || c.Int("OutputPeriod", n.component) != 1)
          throw UnsupportedError("StatisticsExtraction with period != 1 is not used by these recipes");
        const Value* iv = c.Find("IncludeVarinance");
        if (!iv) iv = c.Find("IncludeVariance");
        include_variance_ = iv ? iv->b : true;
        Ref r = Flatten1(n.desc, name);
        if (r.kind != Ref::kFrameLayer)
          throw UnsupportedError("StatisticsExtraction must read a frame-level layer output");
        s.kind = Sym::kExtraction;
        s.extraction_src = r.index;
      } else if (c.type == "StatisticsPoolingComponent") {
        if (n.desc.kind != Desc::kNode) throw UnsupportedError("StatisticsPooling must read one node");
        Sym e = SymOf(n.desc.node);
        if (e.kind != Sym::kExtraction) throw UnsupportedError("StatisticsPooling must follow StatisticsExtraction");
        if (c.Int("InputPeriod", n.component) != 1)
          throw UnsupportedError("StatisticsPooling with input-period != 1 is not used by these recipes");
        if (c.Int("NumLogCountFeatures", n.component) != 0)
          throw UnsupportedError("num-log-count-features > 0 is not used by these recipes");
        if (plan_->pool_src >= 0 && plan_->pool_src != e.extraction_src)
          throw UnsupportedError("more than one statistics-pooling layer");
        plan_->pool_src = e.extraction_src;
        plan_->pool_left_context = c.Int("LeftContext", n.component);
        plan_->pool_right_context = c.Int("RightContext", n.component);
        plan_->pool_output_stddevs = c.Bool("OutputStddevs", n.component);
        if (plan_->pool_output_stddevs && !include_variance_)
          throw FormatError("output-stddevs=true needs include-variance=true");
        plan_->pool_variance_floor = c.Find("VarianceFloor") ? static_cast<float>(c.Real("VarianceFloor", n.component)) : 1e-10f;
        s.kind = Sym::kPooled;
        s.extraction_src = e.extraction_src;
      } else {
        throw UnsupportedError("component type " + c.type + " (node '" + name + "') is not on the extraction path");
      }
    }
    in_progress_.erase(name);
    sym_memo_[name] = s;
    return s;
  }

  struct FlatTap {
    Ref ref;
    int32_t offset;
  };

  void Flatten(const Desc& d, int32_t offset, std::vector<FlatTap>* out) {
    switch (d.kind) {
      case Desc::kAppend:
        for (const Desc& a : d.args) Flatten(a, offset, out);
        return;
      case Desc::kOffset:
        Flatten(d.args[0], offset + d.value, out);
        return;
      case Desc::kRound:
        // Evaluated at t = 0 only (segment level): Round(d, m) is the identity there.
        Flatten(d.args[0], offset, out);
        return;
      case Desc::kNode:
        out->push_back({Materialize(SymOf(d.node), d.node), offset});
        return;
    }
  }

  Ref Flatten1(const Desc& d, const std::string& who) {
    std::vector<FlatTap> taps;
    Flatten(d, 0, &taps);
    if (taps.size() != 1 || (taps[0].offset != 0 && !taps[0].ref.segment))
      throw UnsupportedError("node '" + who + "' must read exactly one node without offsets");
    return taps[0].ref;
  }

  static void BatchNormScaleOffset(const Component& c, const std::string& cname, std::vector<float>* scale,
                                   std::vector<float>* offset) {
    int32_t dim = c.Int("Dim", cname);
    int32_t block = c.Find("BlockDim") ? c.Int("BlockDim", cname) : dim;
    if (block != dim) throw UnsupportedError("BatchNorm " + cname + ": block-dim != dim is not used by these recipes");
    double count = c.Real("Count", cname);
    if (!(count > 0))
      throw FormatError("BatchNorm " + cname + " has <Count> 0: upstream Kaldi would invent random statistics in "
                        "test mode; refusing (SURVEY.md 8a-3)");
    double eps = c.Find("Epsilon") ? c.Real("Epsilon", cname) : 1e-3;
    double rms = c.Find("TargetRms") ? c.Real("TargetRms", cname) : 1.0;
    const Matrix& mean = c.Mat("StatsMean", cname);
    const Matrix& var = c.Mat("StatsVar", cname);
    if (static_cast<int32_t>(mean.data.size()) != dim || static_cast<int32_t>(var.data.size()) != dim)
      throw FormatError("BatchNorm " + cname + ": stats size != dim");
    scale->resize(dim);
    offset->resize(dim);
    for (int32_t i = 0; i < dim; ++i) {
      // scale = target_rms * (max(var, 0) + eps)^-0.5 ; offset = -mean * scale   (fp32 like upstream)
      float s = static_cast<float>(rms) * std::pow(std::max(var.data[i], 0.0f) + static_cast<float>(eps), -0.5f);
      (*scale)[i] = s;
      (*offset)[i] = -mean.data[i] * s;
    }
  }

  Ref Materialize(const Sym& s, const std::string& node_name) {
    Ref r;
    if (s.kind == Sym::kInput) {
      r.kind = Ref::kInputStream;
      r.index = s.input;
      used_inputs_.insert(s.input);
      return r;
    }
    if (s.kind == Sym::kExtraction)
      throw UnsupportedError("StatisticsExtraction output '" + node_name + "' may only feed StatisticsPooling");
    if (s.kind == Sym::kPooled) {
      r.kind = Ref::kPool;
      r.segment = true;
      return r;
    }
    auto m = layer_memo_.find(s.Key());
    if (m != layer_memo_.end()) return m->second;

    const Node& an = NodeByName(s.affine_node);
    const Component& ac = Comp(an);
    std::vector<FlatTap> taps;
    Flatten(an.desc, 0, &taps);
    bool seg = taps[0].ref.segment;
    for (const FlatTap& t : taps)
      if (t.ref.segment != seg) throw UnsupportedError("Append mixes frame-level and segment-level inputs at " + s.affine_node);

    Layer L;
    L.name = s.post_bn.empty() ? (s.relu ? s.affine_node + "+relu" : s.affine_node) : s.post_bn;
    const Matrix* W = nullptr;
    const Value* bv = nullptr;
    if (ac.type == "LinearComponent") {
      W = &ac.Mat("Params", an.component);
    } else {
      W = &ac.Mat("LinearParams", an.component);
      bv = ac.Find("BiasParams");
    }
    L.out_dim = W->rows;
    L.in_dim = W->cols;
    L.weight = W->data;
    if (bv) {
      if (bv->kind != Value::kMatrix || static_cast<int32_t>(bv->m.data.size()) != L.out_dim)
        throw FormatError("bad <BiasParams> in " + an.component);
      L.bias = bv->m.data;
    }
    L.relu = s.relu;
    if (!s.pre_bn.empty()) {
      // affine -> BatchNorm [-> ReLU]: fold the per-row scale/offset into W, b (exact)
      std::vector<float> sc, of;
      BatchNormScaleOffset(net_.components.at(s.pre_bn), s.pre_bn, &sc, &of);
      if (static_cast<int32_t>(sc.size()) != L.out_dim) throw FormatError("BatchNorm dim mismatch after " + s.affine_node);
      if (L.bias.empty()) L.bias.assign(L.out_dim, 0.f);
      for (int32_t o = 0; o < L.out_dim; ++o) {
        for (int32_t k = 0; k < L.in_dim; ++k) L.weight[static_cast<size_t>(o) * L.in_dim + k] *= sc[o];
        L.bias[o] = L.bias[o] * sc[o] + of[o];
      }
    }
    if (!s.post_bn.empty()) {
      BatchNormScaleOffset(net_.components.at(s.post_bn), s.post_bn, &L.scale, &L.offset);
      if (static_cast<int32_t>(L.scale.size()) != L.out_dim) throw FormatError("BatchNorm dim mismatch after " + s.affine_node);
    }

    if (!seg) {
      int32_t k = 0;
      for (const FlatTap& t : taps) {
        Tap tp;
        tp.offset = t.offset;
        if (t.ref.kind == Ref::kInputStream) {
          tp.src = -1 - t.ref.index;
          k += net_.nodes[net_.node_index.at(inputs_[t.ref.index])].dim;
        } else {
          tp.src = t.ref.index;
          k += plan_->frame[t.ref.index].out_dim;
        }
        L.taps.push_back(tp);
      }
      if (k != L.in_dim)
        throw FormatError("layer " + s.affine_node + ": input descriptor has dim " + std::to_string(k) +
                          " but the affine expects " + std::to_string(L.in_dim));
      r.kind = Ref::kFrameLayer;
      r.index = static_cast<int32_t>(plan_->frame.size());
      plan_->frame.push_back(std::move(L));
    } else {
      if (taps.size() != 1) throw UnsupportedError("segment-level Append at " + s.affine_node);
      const Ref& src = taps[0].ref;
      int32_t expect_prev = static_cast<int32_t>(plan_->segment.size()) - 1;
      if (!((src.kind == Ref::kPool && plan_->segment.empty()) ||
            (src.kind == Ref::kSegLayer && src.index == expect_prev)))
        throw UnsupportedError("segment-level layers must form a chain (at " + s.affine_node + ")");
      r.kind = Ref::kSegLayer;
      r.segment = true;
      r.index = static_cast<int32_t>(plan_->segment.size());
      plan_->segment.push_back(std::move(L));
    }
    layer_memo_[s.Key()] = r;
    return r;
  }

  const Nnet& net_;
  Plan* plan_;
  std::vector<std::string> inputs_;
  std::set<int32_t> used_inputs_;
  std::map<std::string, Sym> sym_memo_;
  std::set<std::string> in_progress_;
  std::map<std::string, Ref> layer_memo_;
  bool include_variance_ = true;
  bool frame_output_ = false;
};

}  // namespace

Plan LoadNnet3Plan(const std::string& rxfilename, const std::string& output_node, bool frame_output) {
  Nnet net = ReadNnet(rxfilename);
  if (!output_node.empty()) {
    Node nd;
    nd.kind = Node::kOutput;
    nd.name = "output";
    nd.desc = DescParser(output_node).Parse();
    auto it = net.node_index.find("output");
    if (it != net.node_index.end()) net.nodes[it->second] = nd;
    else {
      net.node_index["output"] = net.nodes.size();
      net.nodes.push_back(nd);
    }
  }
  Plan plan;
  Lowerer(net, &plan, frame_output).Run();
  plan.Finalize();
  return plan;
}

}  // namespace cvx
