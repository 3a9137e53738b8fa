// Host-side counterpart of the reference's LoadEmbed custom op (ops/src/load_dense_embed.cc:28-156):
// rebuild this worker's shard of a row-sharded table from the raw fp32 part files a previous run wrote
// with a possibly different number of workers.
//
//   <ckpt_path>-embedding/<var_name>-part-<p>.bin   rows of old worker p, row j = global row j * P + p
//
// No CUDA here: the op runs on the CPU in the reference too; the caller copies the shard to its arena.
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace er {

// "...-part-<p>.bin" -> p, or -1
static int embed_part_id(const std::string& path) {
  if (path.size() < 5) return -1;
  const size_t dash = path.rfind('-', path.size() - 5);
  if (dash == std::string::npos) return -1;
  const char* tok = path.c_str() + dash + 1;
  if (*tok < '0' || *tok > '9') return -1;
  return std::atoi(tok);
}

}  // namespace er

extern "C" int er_load_embed(const char* ckpt_path, const char* var_name, int32_t task_index,
                             int32_t task_num, int32_t embed_dim, int64_t embed_part_size, float* vals,
                             int64_t* rows_loaded) {
  using namespace er;
  ER_REQUIRE(ckpt_path && var_name && vals, "null argument");
  ER_REQUIRE(task_num > 0 && task_index >= 0 && task_index < task_num, "bad task_index / task_num");
  ER_REQUIRE(embed_dim > 0 && embed_part_size > 0, "bad embed_dim / embed_part_size");
  const std::string folder = std::string(ckpt_path) + "-embedding/";
  const std::string prefix = std::string(var_name) + "-part-";
  DIR* dir = opendir(folder.c_str());
  if (!dir) return fail(ER_ERR_INVALID_ARG, "er_load_embed: cannot open " + folder);
  std::vector<std::string> files;
  while (struct dirent* ent = readdir(dir)) {
    const std::string name = ent->d_name;
    // exactly this variable's parts: "<var>-part-<digits>.bin" (a slot variable "<var>/Adagrad" has its own prefix)
    if (name.compare(0, prefix.size(), prefix) != 0) continue;
    if (name.size() < prefix.size() + 5 || name.compare(name.size() - 4, 4, ".bin") != 0) continue;
    struct stat st;
    const std::string path = folder + name;
    if (stat(path.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) continue;
    if (embed_part_id(path) < 0) continue;
    files.push_back(path);
  }
  closedir(dir);
  if (files.empty()) return fail(ER_ERR_INVALID_ARG, "er_load_embed: no part files " + folder + prefix + "*.bin");
  std::sort(files.begin(), files.end());
  const int64_t parts_old = (int64_t)files.size();
  const int64_t total_rows = embed_part_size * task_num;
  std::memset(vals, 0, sizeof(float) * (size_t)embed_part_size * embed_dim);
  int64_t loaded = 0;
  std::vector<float> buf;
  for (const std::string& path : files) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return fail(ER_ERR_INVALID_ARG, "er_load_embed: cannot read " + path);
    std::fseek(f, 0, SEEK_END);
    const long bytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    const size_t n_flt = (size_t)bytes / sizeof(float);
    buf.resize(n_flt);
    const size_t got = std::fread(buf.data(), sizeof(float), n_flt, f);
    std::fclose(f);
    if (got != n_flt) return fail(ER_ERR_INVALID_ARG, "er_load_embed: short read on " + path);
    const int64_t part_old = embed_part_id(path);
    const int64_t rows_old = (int64_t)(n_flt / (size_t)embed_dim);
    for (int64_t j = 0; j < rows_old; ++j) {
      const int64_t global = j * parts_old + part_old;
      if (global % task_num != task_index || global >= total_rows) continue;
      std::memcpy(vals + (global / task_num) * embed_dim, buf.data() + j * embed_dim, sizeof(float) * embed_dim);
      ++loaded;
    }
  }
  if (rows_loaded) *rows_loaded = loaded;
  // the op's own consistency rule (load_dense_embed.cc:121-127): at most the last local row may be missing
  if (loaded != embed_part_size && loaded + 1 != embed_part_size)
    return fail(ER_ERR_INVALID_ARG, "er_load_embed: part_update_cnt or part_update_cnt + 1 should be equal to "
                                    "embed_part_size, but are: " + std::to_string(loaded) + " and " +
                                    std::to_string(embed_part_size));
  return ER_OK;
}
