// Drop-in drivers for the vector programs the reference's back-end runs right after extraction
// (SURVEY 8f-2), with upstream's argument shapes:
//   egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:102-108
//     ivector-mean ark:$data/spk2utt scp:xvector.scp ark,scp:spk_xvector.ark,spk_xvector.scp ark,t:num_utts.ark
//   egs/sre/v5/run_sre10.sh:107-129
//     ivector-mean scp:xvector.scp mean.vec
//     ivector-subtract-global-mean [mean.vec] scp:xvector.scp ark:- | transform-vec transform.mat ark:- ark:- |
//     ivector-normalize-length ark:- ark:-
// The per-vector arithmetic (subtract, transform, normalise) runs on the device through the C ABI
// (cvx_transform_vectors); grouping by speaker and the global mean are host-side table work.
// nnet3-xvector-compute can also apply the whole chain itself (--mean-vec / --transform-mat /
// --normalize-length), which removes these processes from the pipe altogether.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>

#include "../../../include/cvx.h"
#include "../kaldi_io.h"
#include "extract_main.h"

namespace cvx {
namespace {

int DeviceFromEnv() {
  const char* e = std::getenv("CVX_DEVICE");
  return e ? std::atoi(e) : 0;
}

struct VectorTable {
  std::vector<std::string> keys;
  std::vector<float> data;  // [n, dim]
  int32_t dim = 0;
  int64_t size() const { return static_cast<int64_t>(keys.size()); }
};

// Reads a whole table of vectors; entries of a different dimension are an error like upstream.
VectorTable ReadVectorTable(const std::string& rspecifier) {
  VectorTable t;
  SequentialMatrixReader reader(rspecifier, /*allow_vectors=*/true);
  for (; !reader.Done(); reader.Next()) {
    const Matrix& m = reader.Value();
    if (m.rows != 1) throw KaldiError("entry " + reader.Key() + " of " + rspecifier + " is not a vector");
    if (t.dim == 0) t.dim = m.cols;
    if (m.cols != t.dim)
      throw KaldiError("vector " + reader.Key() + " has dim " + std::to_string(m.cols) + ", expected " + std::to_string(t.dim));
    t.keys.push_back(reader.Key());
    t.data.insert(t.data.end(), m.Data(), m.Data() + m.cols);
  }
  int st = reader.Close();
  if (st != 0) CVX_WARN("the input pipe of " << rspecifier << " exited with status " << st);
  return t;
}

void RunChain(const VectorTable& t, const cvx_vector_chain& chain, std::vector<float>* out, std::vector<float>* ratios) {
  const int32_t od = chain.transform ? chain.transform_rows : t.dim;
  out->assign(static_cast<size_t>(t.size()) * od, 0.f);
  ratios->assign(static_cast<size_t>(t.size()), 0.f);
  if (t.size() == 0) return;
  if (cvx_transform_vectors(DeviceFromEnv(), t.data.data(), t.size(), t.dim, &chain, out->data(), ratios->data()) != CVX_OK)
    throw KaldiError(std::string("cvx_transform_vectors failed: ") + cvx_last_error());
}

void WriteTable(const std::string& wspecifier, const VectorTable& t, const std::vector<float>& data, int32_t dim) {
  VectorWriter writer(wspecifier);
  for (int64_t i = 0; i < t.size(); ++i) writer.Write(t.keys[i], data.data() + static_cast<size_t>(i) * dim, dim);
  writer.Close();
}

int Fail(const std::exception& e) {
  LogMessage("ERROR", "main", __FILE__, __LINE__, e.what());
  return -1 & 0xff;
}

}  // namespace

std::vector<float> ReadVectorFile(const std::string& rxfilename) {
  Input in(rxfilename);
  bool binary = ReadBinaryHeader(in);
  std::vector<float> v;
  ReadVector(in, binary, &v);
  in.Close();
  return v;
}

Matrix ReadMatrixFile(const std::string& rxfilename) {
  Input in(rxfilename);
  bool binary = ReadBinaryHeader(in);
  Matrix m;
  ReadMatrix(in, binary, &m);
  in.Close();
  return m;
}

int IvectorMeanMain(int argc, const char* const* argv) {
  try {
    ParseOptions po(
        "With 3 or 4 arguments, averages vectors over speakers; with 2 arguments, averages all\n"
        "vectors of a table into one vector.\n"
        "Usage: ivector-mean <spk2utt-rspecifier> <ivector-rspecifier> <ivector-wspecifier> [<num-utt-wspecifier>]\n"
        "or: ivector-mean <ivector-rspecifier> <mean-wxfilename>\n"
        "e.g.: ivector-mean data/spk2utt exp/ivectors.ark exp/spk_ivectors.ark exp/spk_num_utts.ark\n"
        "or: ivector-mean exp/ivectors.ark exp/mean.vec\n");
    // upstream: --binary-write, default false (mean.vec is a text vector)
    bool binary_write = false, binary_alias = false;
    po.Register("binary-write", &binary_write, "If true, write the mean vector in binary (2-argument form)");
    po.Register("binary", &binary_alias, "alias of --binary-write");
    po.Read(argc, argv);
    binary_write = binary_write || binary_alias;
    if (po.NumArgs() < 2 || po.NumArgs() > 4) {
      po.PrintUsage();
      return 1;
    }
    if (po.NumArgs() == 2) {
      VectorTable t = ReadVectorTable(po.GetArg(1));
      if (t.size() == 0) throw KaldiError("no vectors read from " + po.GetArg(1));
      std::vector<double> sum(t.dim, 0.0);
      for (int64_t i = 0; i < t.size(); ++i)
        for (int32_t c = 0; c < t.dim; ++c) sum[c] += t.data[static_cast<size_t>(i) * t.dim + c];
      std::vector<float> mean(t.dim);
      for (int32_t c = 0; c < t.dim; ++c) mean[c] = static_cast<float>(sum[c] / static_cast<double>(t.size()));
      Output out;
      out.Open(po.GetArg(2));
      if (binary_write) out.Write("\0B", 2);
      WriteVector(out, binary_write, mean.data(), t.dim);
      if (out.Close() != 0) throw KaldiError("error closing " + po.GetArg(2));
      CVX_LOG("Read " << t.size() << " iVectors; mean written to " << po.GetArg(2));
      return 0;
    }
    // speaker means: random access over the utterance table, sequential over spk2utt
    VectorTable t = ReadVectorTable(po.GetArg(2));
    std::map<std::string, int64_t> index;
    for (int64_t i = 0; i < t.size(); ++i) index[t.keys[i]] = i;
    TableSpec s2u = ParseRspecifier(po.GetArg(1));
    if (s2u.is_scp) throw KaldiError("spk2utt must be an archive (ark:) of text lines");
    Input in(s2u.ark_name);
    VectorWriter writer(po.GetArg(3));
    std::unique_ptr<Output> num_out;
    TableSpec num_spec;
    if (po.NumArgs() == 4) {
      num_spec = ParseWspecifier(po.GetArg(4));
      if (!num_spec.is_ark || num_spec.is_scp) throw KaldiError("the num-utts table must be an archive: " + po.GetArg(4));
      num_out.reset(new Output);
      num_out->Open(num_spec.ark_name);
    }
    int64_t num_spk_done = 0, num_spk_err = 0, num_utt_done = 0, num_utt_err = 0;
    std::vector<double> sum;
    std::vector<float> mean;
    while (!in.Eof()) {
      std::istringstream line(in.ReadLine());
      std::string spk, utt;
      if (!(line >> spk)) continue;
      sum.assign(t.dim, 0.0);
      int32_t n = 0;
      while (line >> utt) {
        auto it = index.find(utt);
        if (it == index.end()) {
          CVX_WARN("No iVector present in input for utterance " << utt);
          ++num_utt_err;
          continue;
        }
        for (int32_t c = 0; c < t.dim; ++c) sum[c] += t.data[static_cast<size_t>(it->second) * t.dim + c];
        ++n;
        ++num_utt_done;
      }
      if (n == 0) {
        CVX_WARN("Not producing output for speaker " << spk << " since no utterances had iVectors");
        ++num_spk_err;
        continue;
      }
      mean.resize(t.dim);
      for (int32_t c = 0; c < t.dim; ++c) mean[c] = static_cast<float>(sum[c] / n);
      writer.Write(spk, mean.data(), t.dim);
      if (num_out) {
        if (num_spec.text) {
          num_out->Write(spk + " " + std::to_string(n) + "\n");
        } else {  // binary int32 table entry: key, "\0B", size byte 4, value
          num_out->Write(spk + " ");
          num_out->Write("\0B", 2);
          char sz = 4;
          num_out->Write(&sz, 1);
          num_out->Write(&n, 4);
        }
      }
      ++num_spk_done;
    }
    in.Close();
    writer.Close();
    if (num_out && num_out->Close() != 0) throw KaldiError("error closing " + num_spec.ark_name);
    CVX_LOG("Computed mean of " << num_spk_done << " speakers (" << num_spk_err << " with no utterances), consisting of "
                                << num_utt_done << " utterances (" << num_utt_err << " absent from input).");
    return num_spk_done != 0 ? 0 : 1;
  } catch (const std::exception& e) {
    return Fail(e);
  }
}

int SubtractGlobalMeanMain(int argc, const char* const* argv) {
  try {
    ParseOptions po(
        "Subtracts a mean from every vector of a table: the mean of the table itself (2 arguments)\n"
        "or a mean read from a file (3 arguments).\n"
        "Usage: ivector-subtract-global-mean <ivector-rspecifier> <ivector-wspecifier>\n"
        "or: ivector-subtract-global-mean <mean-rxfilename> <ivector-rspecifier> <ivector-wspecifier>\n"
        "e.g.: ivector-subtract-global-mean scp:ivectors.scp ark:-\n");
    bool subtract_mean = true;
    po.Register("subtract-mean", &subtract_mean, "If false, copy the vectors unchanged");
    po.Read(argc, argv);
    if (po.NumArgs() != 2 && po.NumArgs() != 3) {
      po.PrintUsage();
      return 1;
    }
    const bool from_file = po.NumArgs() == 3;
    VectorTable t = ReadVectorTable(po.GetArg(from_file ? 2 : 1));
    if (t.size() == 0) throw KaldiError("no vectors read");
    std::vector<float> mean;
    if (from_file) {
      mean = ReadVectorFile(po.GetArg(1));
      if (static_cast<int32_t>(mean.size()) != t.dim)
        throw KaldiError("mean has dim " + std::to_string(mean.size()) + ", vectors have dim " + std::to_string(t.dim));
    } else {
      std::vector<double> sum(t.dim, 0.0);
      for (int64_t i = 0; i < t.size(); ++i)
        for (int32_t c = 0; c < t.dim; ++c) sum[c] += t.data[static_cast<size_t>(i) * t.dim + c];
      mean.resize(t.dim);
      for (int32_t c = 0; c < t.dim; ++c) mean[c] = static_cast<float>(sum[c] / static_cast<double>(t.size()));
    }
    cvx_vector_chain chain;
    std::memset(&chain, 0, sizeof(chain));
    chain.mean = subtract_mean ? mean.data() : nullptr;
    std::vector<float> out, ratios;
    RunChain(t, chain, &out, &ratios);
    WriteTable(po.GetArg(from_file ? 3 : 2), t, out, t.dim);
    CVX_LOG("Wrote " << t.size() << " mean-subtracted iVectors");
    return 0;
  } catch (const std::exception& e) {
    return Fail(e);
  }
}

int TransformVecMain(int argc, const char* const* argv) {
  try {
    ParseOptions po(
        "Multiplies every vector of a table by a matrix; a matrix with one more column than the\n"
        "vectors' dimension is applied as an affine transform (last column = offset).\n"
        "Usage: transform-vec [options] <transform-rxfilename> <feats-rspecifier> <feats-wspecifier>\n");
    po.Read(argc, argv);
    if (po.NumArgs() != 3) {
      po.PrintUsage();
      return 1;
    }
    Matrix tf = ReadMatrixFile(po.GetArg(1));
    VectorTable t = ReadVectorTable(po.GetArg(2));
    if (t.size() > 0 && tf.cols != t.dim && tf.cols != t.dim + 1)
      throw KaldiError("Dimension mismatch: input vector has dim " + std::to_string(t.dim) + " and transform has " +
                       std::to_string(tf.cols) + " columns");
    cvx_vector_chain chain;
    std::memset(&chain, 0, sizeof(chain));
    chain.transform = tf.Data();
    chain.transform_rows = tf.rows;
    chain.transform_cols = tf.cols;
    std::vector<float> out, ratios;
    RunChain(t, chain, &out, &ratios);
    WriteTable(po.GetArg(3), t, out, tf.rows);
    CVX_LOG("Applied transform to " << t.size() << " vectors");
    return t.size() != 0 ? 0 : 1;
  } catch (const std::exception& e) {
    return Fail(e);
  }
}

int NormalizeLengthMain(int argc, const char* const* argv) {
  try {
    ParseOptions po(
        "Scales every vector of a table to length sqrt(dim) (or 1 with --scaleup=false).\n"
        "Usage: ivector-normalize-length [options] <ivector-rspecifier> <ivector-wspecifier>\n"
        "e.g.: ivector-normalize-length ark:ivectors.ark ark:normalized_ivectors.ark\n");
    bool normalize = true, scaleup = true;
    po.Register("normalize", &normalize, "If false, only report the average ratio and copy the vectors");
    po.Register("scaleup", &scaleup, "If true, the target length is sqrt(dim); if false, 1");
    po.Read(argc, argv);
    if (po.NumArgs() != 2) {
      po.PrintUsage();
      return 1;
    }
    VectorTable t = ReadVectorTable(po.GetArg(1));
    cvx_vector_chain chain;
    std::memset(&chain, 0, sizeof(chain));
    chain.normalize_length = normalize ? 1 : 0;
    chain.scaleup = scaleup ? 1 : 0;
    std::vector<float> out, ratios;
    RunChain(t, chain, &out, &ratios);
    double sum = 0.0, sumsq = 0.0;
    for (int64_t i = 0; i < t.size(); ++i) {
      if (ratios[i] == 0.f) CVX_WARN("Zero iVector");
      sum += ratios[i];
      sumsq += static_cast<double>(ratios[i]) * ratios[i];
    }
    WriteTable(po.GetArg(2), t, out, t.dim);
    CVX_LOG("Processed " << t.size() << " iVectors.");
    if (t.size() != 0) {
      const double avg = sum / t.size(), var = std::max(0.0, sumsq / t.size() - avg * avg);
      CVX_LOG("Average ratio of iVector to expected length was " << avg << ", standard deviation was " << std::sqrt(var));
    }
    return t.size() != 0 ? 0 : 1;
  } catch (const std::exception& e) {
    return Fail(e);
  }
}

// ivector-plda-scoring (run_sre10.sh:122-129): every trial "<enrolled-speaker> <test-utterance>" gets
// the PLDA log-likelihood ratio of the two vectors; transform, length normalisation and the scoring
// of the whole trial list run on the device (cvx_plda_score).
int PldaScoringMain(int argc, const char* const* argv) {
  try {
    ParseOptions po(
        "Scores trials with a PLDA model: for every line '<key1> <key2>' of the trials file, key1 a\n"
        "vector of the first (enrolment) table, key2 one of the second (test) table, writes\n"
        "'<key1> <key2> <log-likelihood-ratio>'.\n"
        "Usage: ivector-plda-scoring [options] <plda> <train-ivector-rspecifier> <test-ivector-rspecifier>\n"
        " <trials-rxfilename> <scores-wxfilename>\n"
        "e.g.: ivector-plda-scoring --num-utts=ark:exp/train/num_utts.ark plda ark:exp/train/spk_ivectors.ark\n"
        " ark:exp/test/ivectors.ark trials scores\n");
    std::string num_utts_rspecifier;
    bool normalize_length = false, simple_length_norm = false;
    po.Register("num-utts", &num_utts_rspecifier,
                "Table of the number of utterances behind every enrolment vector (written by ivector-mean); default 1");
    po.Register("normalize-length", &normalize_length, "If true, do PLDA length normalisation of the transformed vectors");
    po.Register("simple-length-normalization", &simple_length_norm,
                "If true, normalise to length sqrt(dim) instead of the PLDA-specific normalisation");
    po.Read(argc, argv);
    if (po.NumArgs() != 5) {
      po.PrintUsage();
      return 1;
    }
    // <Plda> mean (vector) transform (matrix) psi (vector) </Plda>: stored in double precision upstream
    // (DV / DM) and kept in double here
    std::vector<double> mean, psi, tf;
    struct { int32_t rows = 0, cols = 0; } tf_f;
    {
      Input in(po.GetArg(1));
      const bool binary = ReadBinaryHeader(in);
      ExpectToken(in, "<Plda>");
      int32_t r = 0, c = 0;
      ReadVectorOrMatrixDouble(in, binary, &r, &c, &mean, nullptr);
      if (r > 1) throw KaldiError("expected the PLDA mean vector in " + po.GetArg(1));
      ReadVectorOrMatrixDouble(in, binary, &tf_f.rows, &tf_f.cols, &tf, nullptr);
      ReadVectorOrMatrixDouble(in, binary, &r, &c, &psi, nullptr);
      if (r > 1) throw KaldiError("expected the PLDA psi vector in " + po.GetArg(1));
      ExpectToken(in, "</Plda>");
      int st = in.Close();
      if (st != 0) throw KaldiError("the PLDA input pipe exited with status " + std::to_string(st));
    }
    if (tf_f.cols != static_cast<int32_t>(mean.size()) || tf_f.rows != static_cast<int32_t>(psi.size()))
      throw KaldiError("inconsistent dimensions in the PLDA model " + po.GetArg(1));

    VectorTable train = ReadVectorTable(po.GetArg(2)), test = ReadVectorTable(po.GetArg(3));
    if ((train.size() && train.dim != tf_f.cols) || (test.size() && test.dim != tf_f.cols))
      throw KaldiError("vector dimension does not match the PLDA model (" + std::to_string(tf_f.cols) + ")");
    std::map<std::string, int32_t> train_index, test_index;
    for (int64_t i = 0; i < train.size(); ++i) train_index[train.keys[i]] = static_cast<int32_t>(i);
    for (int64_t i = 0; i < test.size(); ++i) test_index[test.keys[i]] = static_cast<int32_t>(i);
    std::vector<int32_t> num_utts(static_cast<size_t>(train.size()), 1);
    if (!num_utts_rspecifier.empty()) {
      TableSpec spec = ParseRspecifier(num_utts_rspecifier);
      if (spec.is_scp) throw KaldiError("--num-utts must be an archive of integers (ark:), as written by ivector-mean ... ark,t:");
      Input in(spec.ark_name);
      std::map<std::string, int32_t> seen;
      while (!in.Eof()) {
        std::istringstream line(in.ReadLine());
        std::string key;
        int32_t n = 0;
        if (!(line >> key)) continue;
        if (!(line >> n) || n <= 0) throw KaldiError("bad line for " + key + " in " + num_utts_rspecifier + " (text archives only)");
        seen[key] = n;
      }
      in.Close();
      for (int64_t i = 0; i < train.size(); ++i) {
        auto it = seen.find(train.keys[i]);
        if (it == seen.end()) throw KaldiError("Number of utterances not given for speaker " + train.keys[i]);
        num_utts[i] = it->second;
      }
    }

    std::vector<int32_t> ti, tj;
    std::vector<std::pair<std::string, std::string>> names;
    int64_t num_err = 0;
    {
      Input in(po.GetArg(4));
      while (!in.Eof()) {
        std::istringstream line(in.ReadLine());
        std::string k1, k2;
        if (!(line >> k1)) continue;
        if (!(line >> k2)) throw KaldiError("Bad line in the trials file: expected two or more fields");
        auto a = train_index.find(k1);
        auto b = test_index.find(k2);
        if (a == train_index.end()) {
          CVX_WARN("Key " << k1 << " not present in training iVectors.");
          ++num_err;
          continue;
        }
        if (b == test_index.end()) {
          CVX_WARN("Key " << k2 << " not present in test iVectors.");
          ++num_err;
          continue;
        }
        ti.push_back(a->second);
        tj.push_back(b->second);
        names.emplace_back(k1, k2);
      }
      in.Close();
    }
    std::vector<float> scores(ti.size(), 0.f);
    if (!ti.empty()) {
      cvx_plda model;
      std::memset(&model, 0, sizeof(model));
      model.mean = mean.data();
      model.transform = tf.data();
      model.psi = psi.data();
      model.dim_in = tf_f.cols;
      model.dim = tf_f.rows;
      model.normalize_length = normalize_length ? 1 : 0;
      model.simple_length_norm = simple_length_norm ? 1 : 0;
      if (cvx_plda_score(DeviceFromEnv(), &model, train.data.data(), num_utts.data(), static_cast<int32_t>(train.size()),
                         test.data.data(), static_cast<int32_t>(test.size()), ti.data(), tj.data(),
                         static_cast<int64_t>(ti.size()), scores.data()) != CVX_OK)
        throw KaldiError(std::string("cvx_plda_score failed: ") + cvx_last_error());
    }
    Output out;
    out.Open(po.GetArg(5));
    std::string buf;
    double sum = 0.0, sumsq = 0.0;
    for (size_t t = 0; t < names.size(); ++t) {
      char num[48];
      std::snprintf(num, sizeof num, "%.9g", scores[t]);
      buf += names[t].first + " " + names[t].second + " " + num + "\n";
      sum += scores[t];
      sumsq += static_cast<double>(scores[t]) * scores[t];
      if (buf.size() > (1u << 20)) {
        out.Write(buf);
        buf.clear();
      }
    }
    out.Write(buf);
    if (out.Close() != 0) throw KaldiError("error closing " + po.GetArg(5));
    if (!names.empty()) {
      const double m = sum / names.size(), sd = std::sqrt(std::max(0.0, sumsq / names.size() - m * m));
      CVX_LOG("Mean score was " << m << ", standard deviation was " << sd);
    }
    CVX_LOG("Processed " << names.size() << " trials, " << num_err << " had errors.");
    return !names.empty() ? 0 : 1;
  } catch (const std::exception& e) {
    return Fail(e);
  }
}

}  // namespace cvx
