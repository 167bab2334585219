/* avsr_io.h - C ABI of the input pipeline in front of the hot path.
 *
 * The reference reads TFRecord files of tf.train.SequenceExample protos with tf.data
 * (avsr/io_utils.py:21-165 one stream, :168-257 two streams) that avsr/dataset_writer.py wrote
 * (labels :290-311, features :439-458, video :461-498).  TensorFlow is not a dependency here: this library
 * reads and writes the same bytes - the TFRecord framing
 *     u64 length | u32 masked_crc32c(length) | bytes | u32 masked_crc32c(bytes)       (little endian)
 * and the protobuf wire format of SequenceExample{context = 1, feature_lists = 2} - and assembles zero-padded
 * batches (tf.data padded_batch, io_utils.py:111-122) straight into caller-provided (pinned) host buffers with a
 * small thread pool (the reference maps with num_parallel_calls = 4, io_utils.py:88-98).
 *
 * Host-only code (no CUDA).  All pointers are HOST pointers.  Return 0 on success, otherwise
 * avsr_io_last_error() describes the failure.
 */
#ifndef AVSR_IO_H_
#define AVSR_IO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* avsr_io_last_error(void);

/* crc32c (Castagnoli) and TFRecord's masking ((crc >> 15 | crc << 17) + 0xa282ead8) */
uint32_t avsr_io_crc32c(const void* data, size_t n);
uint32_t avsr_io_masked_crc32c(const void* data, size_t n);

/* ---- reading ------------------------------------------------------------------------------------ */
typedef struct AvsrIoFile AvsrIoFile;

enum { AVSR_IO_FEATURE = 0, AVSR_IO_VIDEO = 1, AVSR_IO_LABELS = 2 };

typedef struct AvsrIoInfo {
  int kind;            /* AVSR_IO_FEATURE: context has input_size; AVSR_IO_VIDEO: width/height/channels; AVSR_IO_LABELS */
  int has_aus;         /* feature_lists has 'aus' (io_utils.py:334-340) */
  long long n_records;
  long long feat;      /* values per step: input_size, or width*height*channels; 1 for labels */
  int width, height, channels; /* video only (io_utils.py:318-332: channels defaults to 1) */
  char unit[32];       /* labels only: the context's 'unit' string */
} AvsrIoInfo;

/* Opens a TFRecord file, indexes every record (verifying the length CRCs; verify_data != 0 also checks every
 * payload CRC) and inspects the first example like _get_input_shape_from_record (io_utils.py:308-341). */
int avsr_io_open(const char* path, int verify_data, AvsrIoFile** out);
void avsr_io_close(AvsrIoFile* f);
int avsr_io_info(const AvsrIoFile* f, AvsrIoInfo* info);
/* per record: number of steps (context input_length / labels_length) and the context filename (NUL-terminated copy,
 * truncated to cap - 1 bytes) */
int avsr_io_lengths(AvsrIoFile* f, long long* lengths /*[n_records]*/);
int avsr_io_filename(AvsrIoFile* f, long long idx, char* dst, int cap);

/* padded_batch of records idx[0..n): dst[n, t_pad, feat] zero-padded floats, lens[n]; aus_dst[n, t_pad, 2] or NULL.
 * reverse != 0 reverses the valid steps of each row (reverse_input, io_utils.py:104-107: tf.reverse on the unpadded
 * example).  Records are decoded by up to n_threads workers. */
int avsr_io_fill_inputs(AvsrIoFile* f, const long long* idx, int n, int t_pad, float* dst, float* aus_dst,
                        int32_t* lens, int reverse, int n_threads);
/* labels of records idx[0..n): dst[n, l_pad] int32 zero-padded, with `eos` appended after the stored labels
 * (io_utils.py:80-83), lens[n] = stored length + 1. */
int avsr_io_fill_labels(AvsrIoFile* f, const long long* idx, int n, int l_pad, int32_t eos, int32_t* dst,
                        int32_t* lens);

/* ---- writing (dataset_writer.py's examples; used by the synthetic-data generator and the tests) -- */
typedef struct AvsrIoWriter AvsrIoWriter;
int avsr_io_writer_open(const char* path, AvsrIoWriter** out);
int avsr_io_writer_close(AvsrIoWriter* w);
/* make_feature_example (dataset_writer.py:439-458): inputs [steps, size] */
int avsr_io_write_feature(AvsrIoWriter* w, const char* sentence_id, const float* inputs, int steps, int size);
/* make_video_example (dataset_writer.py:461-498): frames [steps, height, width, channels]; aus [steps, 2] or NULL */
int avsr_io_write_video(AvsrIoWriter* w, const char* sentence_id, const float* frames, int steps, int height,
                        int width, int channels, const float* aus);
/* _make_label_example (dataset_writer.py:290-311): labels [n] without EOS */
int avsr_io_write_labels(AvsrIoWriter* w, const char* label_id, const int64_t* labels, int n, const char* unit);

#ifdef __cplusplus
}
#endif
#endif /* AVSR_IO_H_ */
