/*
 * gpet_b200.h -- C ABI of libgpet_b200.so, the B200-native (sm_100a) implementation of gPET's
 * Monte-Carlo hot path (source sampling -> phantom transport -> detector transport -> digitizer).
 *
 * The reference (utaresearch/gPET) has no FFI/plugin layer: its surface is the CLI `./gPET input_PET.in`,
 * the input/output file formats and the host call sequence declared in gPET.h:128-166.  Every entry point
 * below names the reference host function (file:line under /root/reference) whose job it takes over.
 * Only plain pointers, sizes and POD structs cross this boundary (no C++/torch types).
 *
 * Conventions
 *   - every function returns GPET_OK (0) or a negative gpet_status; gpet_last_error() gives the text.
 *     The library never calls exit() (the reference's CUDA_CALL/FILEEXIST macros do, gPET.h:17-18).
 *   - a context is single-threaded and owns one GPU (reference: one device, default stream, main.cu:54,192).
 *   - device = -1 creates a HOST-ONLY context: loaders/parsers/getters work, every compute entry point fails
 *     with GPET_ERR_NO_DEVICE.  There is no CPU fallback for the hot path.
 *   - units follow the reference: length cm, energy eV, time us (fp64), c = 29979.2458 cm/us.
 */
#ifndef GPET_B200_H
#define GPET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPET_ABI_VERSION 5  /* 5: gpet_set_singles_format, gpet_result_singles_compact; 4: 64-bit history numbers (gpet_set_first_pair), gpet_peek_config_device, results of large file runs streamed; 3: see below; 2: gpet_transport_params + record_psf / record_sphere, gpet_digitizer_params + noise_*, gpet_set_psf_output, gpet_stage_noise
                             * 3: coincidence classes: gpet_digitizer_params + coinc_pair_shift, gpet_stats + trues / scatters / randoms,
                             *    gpet_fetch_coincidence_classes, gpet_result_coincidence_classes, gpet_mark_scattered */

typedef enum gpet_status {
    GPET_OK = 0,
    GPET_ERR_ARG = -1,        /* bad argument / state (e.g. stage called before its inputs were loaded) */
    GPET_ERR_IO = -2,         /* file missing / short / malformed */
    GPET_ERR_CUDA = -3,       /* CUDA runtime error (text in gpet_last_error) */
    GPET_ERR_NO_DEVICE = -4,  /* compute call on a host-only context */
    GPET_ERR_CAPACITY = -5,   /* a device buffer overflowed (hits / events); rerun with larger capacity */
    GPET_ERR_FORMAT = -6      /* table/geometry content inconsistent (e.g. material id out of range) */
} gpet_status;

typedef struct gpet_ctx gpet_ctx;

/* ---- records ------------------------------------------------------------------------------------------ */

/* 48-byte single/event record, byte-identical to the reference `Event` (gPET.h:87-92) and to what
 * output/readOutput.m:18-34 reads: 6 x int32, 1 x float64, 4 x float32. */
typedef struct gpet_event {
    int32_t parn, pann, modn, cryn, siten, eventid;
    double  t;          /* us */
    float   E, x, y, z; /* eV; panel-local cm */
} gpet_event;

/* 96-byte coincidence record (extension: the reference has no coincidence sorter):
 * the two singles of the pair, earlier one first. */
typedef struct gpet_coincidence {
    gpet_event a, b;
} gpet_coincidence;

/* One detector hit = one row of HitsID.dat (5 x int32) + one row of Hits.dat (5 x float32)
 * (gPET_kernals.cu:1062-1080, readOutput.m:3-16).  t is also carried in fp64 (extension). */
typedef struct gpet_hit {
    int32_t parn, pann, modn, cryn, type; /* type: 1 Compton, 2 Compton below Eabs, 3 Rayleigh, 4 photoelectric */
    float   E, t32, x, y, z;
    double  t;
} gpet_hit;

/* Photon phase-space record (host side), same quantities as the reference's x/vx/d_time/d_eventid
 * arrays (gPET_kernals.h:6-9).  t <= 0 marks a dead/empty photon. */
typedef struct gpet_photon {
    float   x, y, z, E;
    float   vx, vy, vz;
    int32_t nscat;      /* phantom scatter count (extension) */
    double  t;
    int32_t eventid, parn;
} gpet_photon;

/* Panel geometry, same fields as the reference `object_t` (gPET.h:54-78). */
typedef struct gpet_panel {
    int32_t panel;
    float lengthx, lengthy, lengthz;
    float MODx, MODy, MODz;
    float Mspacex, Mspacey, Mspacez;
    float LSOx, LSOy, LSOz;
    float spacex, spacey, spacez;
    float offsetx, offsety, offsetz;
    float directionx, directiony, directionz;
    float UniXx, UniXy, UniXz;
    float UniYx, UniYy, UniYz;
    float UniZx, UniZy, UniZz;
} gpet_panel;

#define GPET_MAX_SURFACES 5  /* constants.h:19 MAXSURFACE */
#define GPET_MAX_SOURCES 64
#define GPET_MAX_ISOTOPES 16
#define GPET_MAX_MATERIALS 16

/* Digitizer parameters = input_PET.in fields 18-22 (main.cu:149-182) + extensions. */
typedef struct gpet_digitizer_params {
    int32_t readout_depth;   /* 0 world, 1 panel, 2 module, 3 crystal (gPET_kernals.cu:756-784) */
    int32_t readout_policy;  /* 0 winner-take-all, 1 energy centroid (forces depth 2) */
    float   threshold_eV;    /* Eth: first energy window is [Eth, 2e6] (gPET.cu:155,393) */
    int32_t blur_policy;     /* 0: R = sqrt(Eref/E)*Rref ; 1: R = Rref + slope*(E-Eref)/1e6 (gPET_kernals.cu:822-823) */
    float   blur_Eref, blur_Rref, blur_slope, blur_space;
    int32_t dead_level;      /* 0..3; 3 = keep the siten left by readout (gPET.cu:164-169) */
    int32_t dead_type;       /* 0 paralyzable, 1 non-paralyzable (gPET_kernals.cu:657-698) */
    float   dead_time_us;
    float   ewin_min, ewin_max;
    /* --- extensions (no reference counterpart; 0 disables) --- */
    float   time_blur_sigma_us;  /* Gaussian blur of t */
    float   coinc_window_us;     /* coincidence window; 0 = no coincidence sorting */
    int32_t coinc_policy;        /* 0 drop multiples (window with != 2 singles), 1 all pairs with the window opener */
    int32_t coinc_min_panel_diff;/* minimum |panel difference| (cyclic) for a valid pair; 0 = any two distinct sites */
    /* noise singles = the reference's addnoise kernel (gPET_kernals.cu:699-735; declared in gPET.h:191, never launched
     * there, so its four arguments have no input_PET.in field): a Poisson process of mean gap noise_mean_gap_us over the
     * whole detector, drawn per time slice of noise_interval_us, E ~ N(noise_Emean, noise_sigma), uniform crystal.
     * noise_mean_gap_us <= 0 disables it. */
    float   noise_mean_gap_us, noise_Emean_eV, noise_sigma_eV, noise_interval_us;
    /* coincidence classes: two singles belong to the same annihilation iff eventid >> coinc_pair_shift agree.  0 for the
     * isotope source and the positron PSF (both photons carry the pair's eventid, gPET_kernals.cu:546-547, 590-591);
     * 1 for a photon PSF whose records 2k, 2k+1 are the two photons of pair k (eventid = record index, initialize.cu:103). */
    int32_t coinc_pair_shift;
} gpet_digitizer_params;

/* Transport parameters = input_PET.in fields 2, 12, 15, 17 (main.cu:57-60, 117-120, 135-147). */
typedef struct gpet_transport_params {
    float   noncollinearity_rad; /* sigma of the Gaussian acollinearity */
    int32_t use_positron_range;
    float   eabs_eV;             /* photon absorption energy */
    int32_t nsurface;
    float   surface[10 * GPET_MAX_SURFACES];
    int32_t record_hits;         /* OUTPUTHIT (constants.h:6): keep Hits/HitsID rows */
    int32_t record_psf;          /* 1 = the RECORDPSF == -1 branch of photon() (gPET_kernals.cu:288-294, getDistance :148-171):
                                    a photon that leaves the phantom is moved onto the recording sphere */
    float   record_sphere[4];    /* centre x y z and radius, cm (input_PET.in field 14, main.cu:127-133) */
} gpet_transport_params;

/* Run counters (the numbers the reference prints per epoch, gPET.cu:293,364,382,398,415,423). */
typedef struct gpet_stats {
    uint64_t pairs;            /* annihilation pairs emitted */
    uint64_t photons_phantom_out; /* photons alive after the phantom */
    uint64_t photons_on_panel; /* photons that entered a panel front face */
    uint64_t hits;             /* rows of Hits.dat */
    uint64_t events_adder;     /* "counts of events after adder" */
    uint64_t events_threshold; /* "after thresholder" */
    uint64_t events_deadtime;  /* "after deadtime" */
    uint64_t singles;          /* "counts of singles" */
    uint64_t coincidences;     /* extension */
    uint64_t overflow_hits, overflow_events, overflow_adder; /* dropped records (must be 0) */
    uint64_t frames;
    uint64_t kernel_launches;  /* launches of this library's own kernels */
    double   ms_source, ms_phantom, ms_detector, ms_digitizer, ms_total; /* CUDA-event times */
    /* coincidence classes (extension): trues + scatters + randoms == coincidences */
    uint64_t trues;            /* same annihilation, neither photon interacted in the phantom */
    uint64_t scatters;         /* same annihilation, at least one photon Compton- or Rayleigh-scattered in the phantom */
    uint64_t randoms;          /* different annihilations, or a noise single */
} gpet_stats;

/* ---- lifecycle ---------------------------------------------------------------------------------------- */

int gpet_abi_version(void);
/* iniDevice(deviceNo) (iniDevice.cu:42-58).  device < 0: host-only context. */
int gpet_create(int device, gpet_ctx** out);
void gpet_destroy(gpet_ctx* ctx);
const char* gpet_last_error(const gpet_ctx* ctx);
/* Run all kernels on an external cudaStream_t (e.g. torch's current stream). NULL = library-owned stream. */
int gpet_set_stream(gpet_ctx* ctx, void* cuda_stream);
/* Philox key.  Replaces srand(time(NULL)) + curand_init (initialize.cu:256-272). */
int gpet_set_seed(gpet_ctx* ctx, uint64_t seed);
/* Capacities in photons/hits/events per frame (reference: NPART, NSSTACK/5, 3*NPART; constants.h:13-16). */
int gpet_set_capacity(gpet_ctx* ctx, uint64_t max_photons, uint64_t max_hits, uint64_t max_events);

/* ---- loaders: the reference's input contract ------------------------------------------------------------ */

/* main.cu:50-184 positional parse of input_PET.in + the whole init chain (main.cu:192-229): loads the phantom,
 * geometry and source/PSF files it names (relative to base_dir, NULL = cwd) and the table set
 * `<data_dir>/input4gPET.*` (NULL = "<base_dir>/data"). */
int gpet_load_config_file(gpet_ctx* ctx, const char* input_file, const char* base_dir, const char* data_dir);
/* rmater/rlamph/rcompt/rcmpsf/rphote/rrayle/rrayff (initialize.cu:279-748).  prefix e.g. "data/input4gPET".
 * Table dimensions come from the file headers.  A prefix ending in ".gpettab" loads a packed binary set. */
int gpet_load_tables(gpet_ctx* ctx, const char* prefix);
int gpet_save_tables_packed(gpet_ctx* ctx, const char* path);
/* loadPhantom + initPhantom + iniwck(phantom) (initialize.cu:32-74, 831-884, 773-829). int32 mat, float32 density, x fastest. */
int gpet_load_phantom_files(gpet_ctx* ctx, const char* mat_file, const char* den_file,
                            const int32_t dim[3], const float offset[3], const float size[3]);
int gpet_set_phantom(gpet_ctx* ctx, const int32_t* mat, const float* dens,
                     const int32_t dim[3], const float offset[3], const float size[3]);
/* read_file_ro + iniPanel + iniwck(detector) (detector.cu:64-285, initialize.cu:969-1224, 919-966). */
int gpet_load_geometry(gpet_ctx* ctx, const char* geo_file);
/* loadIsotopes / readSource (initialize.cu:10-31, 116-144). */
int gpet_load_isotopes(gpet_ctx* ctx, const char* isotopes_file);
int gpet_load_source(gpet_ctx* ctx, const char* source_file);
/* readParticle (initialize.cu:76-115): 8 x fp64 per record (x y z t vx vy vz E). ptype 0 positron, 1 photon. */
int gpet_load_psf(gpet_ctx* ctx, const char* psf_file, int64_t max_particles, int ptype);
int gpet_set_digitizer(gpet_ctx* ctx, const gpet_digitizer_params* p);
int gpet_get_digitizer(const gpet_ctx* ctx, gpet_digitizer_params* p);
int gpet_set_transport(gpet_ctx* ctx, const gpet_transport_params* p);
int gpet_get_transport(const gpet_ctx* ctx, gpet_transport_params* p);
/* tstart, tend in seconds (input_PET.in field 13). */
int gpet_set_time_window(gpet_ctx* ctx, float tstart_s, float tend_s);
/* Override the atom count of one source (synthetic scaling of the activity). */
int gpet_set_source_atoms(gpet_ctx* ctx, int source_index, uint64_t natom);

/* ---- getters (host state; used by the parity tests) --------------------------------------------------- */

int gpet_get_num_panels(const gpet_ctx* ctx);
int gpet_get_panels(const gpet_ctx* ctx, gpet_panel* out, int cap);
/* moduleNy, crystalNy, moduleN, crystalN (initialize.cu:1074-1086) and the two panel materials/densities. */
int gpet_get_geometry_counts(const gpet_ctx* ctx, int32_t counts[4], int32_t mat[2], float dens[2]);
/* nmat, n energies, e0, de of the 1-D tables; surface dims of cmpsf and rayff. */
int gpet_get_table_dims(const gpet_ctx* ctx, int32_t* nmat, int32_t* nen, float* e0, float* e1,
                        int32_t cmpsf_dims[2], float cmpsf_step[2], int32_t rayff_dims[2], float rayff_step[2]);
/* which: 0 lamph, 1 compt, 2 phote, 3 rayle (nmat*nen floats); 4 cmpsf surface, 5 rayff surface (nmat*ncp*ne);
 * 6 phantom majorant Sigma_max(E) (nen floats, 1/cm), 7 detector majorant. Returns number of floats written. */
int64_t gpet_get_table(const gpet_ctx* ctx, int which, float* out, int64_t cap);
int gpet_get_num_sources(const gpet_ctx* ctx);
int gpet_get_source(const gpet_ctx* ctx, int i, uint64_t* natom, int32_t* type, int32_t* shape, float coeff[6]);
int gpet_get_num_isotopes(const gpet_ctx* ctx);
int gpet_get_isotope(const gpet_ctx* ctx, int i, float* halflife, float* ratio, float coef[8]);
int64_t gpet_get_num_psf(const gpet_ctx* ctx);

/* ---- the hot path, stage by stage (device-resident between stages) ------------------------------------- */

/* Frame planning: sampleParticle's epoch loop (gPET.cu:204-208, 260-282) + findT (gPET.cu:439-452), with
 * the per-atom Bernoulli sweep of setPosition replaced by binomial thinning per source.
 * Returns the number of frames planned for [tstart, tend]; max_pairs_per_frame 0 = capacity/2. */
int64_t gpet_plan_frames(gpet_ctx* ctx, uint64_t max_pairs_per_frame);
/* npairs of frame f (after planning). */
int64_t gpet_frame_pairs(const gpet_ctx* ctx, int64_t frame);
/* Full description of frame f: start (s), length (s), global index of its first pair, pairs per source
 * (pairs_per_source must hold gpet_get_num_sources() entries). */
int gpet_get_frame(const gpet_ctx* ctx, int64_t frame, double* t0_s, double* dt_s, uint64_t* first_pair,
                   uint64_t* pairs_per_source);

/* S2/S3/S4/S5 setPosition (gPET_kernals.cu:483-561): sample the pairs of frame f into photon queue 0. */
int gpet_stage_source(gpet_ctx* ctx, int64_t frame);
/* Load PSF photons [first, first+n) into queue 0 (simulateParticle batch upload, gPET.cu:47-61);
 * positron PSF (ptype 0) goes through setPositionForPhoton (gPET_kernals.cu:563-604). */
int gpet_stage_psf(gpet_ctx* ctx, int64_t first, int64_t n);
/* P1 photon (gPET_kernals.cu:256-345): queue 0 -> queue 1 (photons alive after the phantom). */
int gpet_stage_phantom(gpet_ctx* ctx);
/* X1 photonde + D1 adder + D2 readout (gPET_kernals.cu:839-1233, 737-813): queue 1 -> hits + events. */
int gpet_stage_detector(gpet_ctx* ctx);
/* The three stages above up to the panel face in ONE kernel (what gpet_run uses): frame >= 0 samples that frame's pairs
 * (setPosition), frame = -1 starts from queue 0 (PSF batches); then photon (gPET_kernals.cu:256-345) and the panel-entry
 * prologue of photonde (:963-1009) with the photon state in registers; only photons that entered a panel are written,
 * to queue 2 (panel-local frame).  Same Philox counters as the staged calls, hence the same photons; queues 0 and 1
 * are not materialised (their counters still hold the tallies). */
int gpet_stage_front(gpet_ctx* ctx, int64_t frame);
/* X1 photonde after its panel-entry prologue + D1 adder + D2 readout: queue 2 -> hits + events. */
int gpet_stage_panel_transport(gpet_ctx* ctx);
/* D3-D7 blur, energywindow, sort by t, setSitenum, orderevents, deadtime, energywindow (gPET.cu:385-424)
 * + the coincidence sorter extension: events -> singles (time sorted) [+ coincidences]. */
int gpet_stage_digitize(gpet_ctx* ctx);
/* addnoise (gPET_kernals.cu:699-735): append the noise events with t_lo_us <= t < t_hi_us to the event buffer (after the
 * detector stage, before gpet_stage_digitize).  gpet_run does this per frame when noise is enabled. */
int gpet_stage_noise(gpet_ctx* ctx, double t_lo_us, double t_hi_us);

/* Host <-> device access to the stage buffers. which_queue: 0 after source, 1 after phantom, 2 entered a panel. */
int64_t gpet_queue_size(gpet_ctx* ctx, int which_queue);
int gpet_put_photons(gpet_ctx* ctx, int which_queue, const gpet_photon* in, int64_t n);
int64_t gpet_fetch_photons(gpet_ctx* ctx, int which_queue, gpet_photon* out, int64_t cap);
int gpet_put_events(gpet_ctx* ctx, const gpet_event* in, int64_t n);
int64_t gpet_fetch_events(gpet_ctx* ctx, gpet_event* out, int64_t cap);    /* post adder/readout ("adder.dat") */
int64_t gpet_fetch_hits(gpet_ctx* ctx, gpet_hit* out, int64_t cap);
int64_t gpet_fetch_singles(gpet_ctx* ctx, gpet_event* out, int64_t cap);   /* "singles.dat" of the last frame */
int64_t gpet_fetch_coincidences(gpet_ctx* ctx, gpet_coincidence* out, int64_t cap);
/* Coincidence classes of the last frame (extension; the reference has no sorter, F2, and no scatter flag,
 * F11): one byte per record of gpet_fetch_coincidences, 0 true / 1 scatter / 2 random.  Random = the two singles stem
 * from different annihilations (eventid >> coinc_pair_shift differ) or one is a noise single (parn == -1); scatter =
 * same annihilation and at least one of the two photons had a Compton or Rayleigh interaction in the phantom
 * (gPET_kernals.cu:304-334) before it entered its panel; true = the rest.  totals (may be NULL) = {trues, scatters,
 * randoms} of the frame, counted over all coincidences even when the record buffer overflowed; out may be NULL. */
int64_t gpet_fetch_coincidence_classes(gpet_ctx* ctx, uint8_t* out, int64_t cap, uint64_t totals[3]);
/* Replay (gpet_put_events -> gpet_stage_digitize) has no transport history: this marks the photons with the given
 * photon numbers (gpet_event.parn) as scattered in the phantom.  Call after gpet_put_events, which forgets all marks;
 * the photon numbers of one event list must be distinct modulo the photon capacity rounded up to a power of two (true
 * for the contiguous numbers the source and PSF stages assign). */
int gpet_mark_scattered(gpet_ctx* ctx, const int32_t* parn, int64_t n);
/* counts[0..3] = events in, after thresholder, after deadtime, singles (gPET.cu:382,398,415,423). */
int gpet_last_counts(gpet_ctx* ctx, uint64_t counts[4]);

/* ---- whole-path entry points --------------------------------------------------------------------------- */

/* The replay entry (bit-exact pin): host list of post-adder events ("adder.dat") -> singles.
 * Copies H2D, runs gpet_stage_digitize, copies D2H. counts may be NULL. */
int gpet_digitize(gpet_ctx* ctx, const gpet_event* in, int64_t n, gpet_event* out, int64_t cap,
                  int64_t* n_out, uint64_t counts[4]);
/* sampleParticle / simulateParticle (gPET.cu:13-437): all frames, source or PSF mode according to the loaded
 * inputs.  Singles (and coincidences) of all frames accumulate in pinned host memory owned by the context;
 * output_dir != NULL additionally appends HitsID.dat/Hits.dat/adder.dat/singles.dat[/coincidences.dat] there
 * with the reference layouts (gPET.cu:367-383, 424). */
int gpet_run(gpet_ctx* ctx, const char* output_dir, gpet_stats* stats);
/* Same, but nothing leaves the device except the counters (bench `value`: inputs resident, no D2H of records). */
int gpet_run_resident(gpet_ctx* ctx, gpet_stats* stats);
int64_t gpet_result_singles(gpet_ctx* ctx, const gpet_event** ptr);
int64_t gpet_result_coincidences(gpet_ctx* ctx, const gpet_coincidence** ptr);
/* How gpet_run brings coincidences to the host (extension, like the sorter itself).  GPET_COINC_RECORDS (default): the
 * two 48-byte singles side by side, 96 B per coincidence.  GPET_COINC_PAIRS: two uint32 indices into the run's singles
 * list (gpet_result_singles), 8 B per coincidence -- the records are a gather the host can do when it needs them:
 * gpet_result_coincidences builds them on demand, and coincidences.dat is written from them, so both formats give
 * the same files and the same records. */
enum { GPET_COINC_RECORDS = 0, GPET_COINC_PAIRS = 1 };
int gpet_set_coincidence_format(gpet_ctx* ctx, int format);
/* How gpet_run(NULL) brings SINGLES to the host (extension; files always hold the reference's 48-byte Event, detector.cu:287-307).
 * GPET_SINGLES_RECORDS (default): 48-byte records.  GPET_SINGLES_COMPACT: 32-byte records that hold the same information --
 * end to end the run is bound by the device-to-host copy of the singles, and a third of an Event is redundant in source mode:
 *   t, E, x, y, z as they are; eventid; ids = pann | modn << 8 | cryn << 20 | (parn & 1) << 31.
 * parn is (eventid << 1 | parn & 1) & 0x7fffffff (two photons per annihilation, device numbering) and siten follows from the
 * dead-time level (or, at level 3, the readout level) and the three ids.  gpet_result_singles expands them on demand, byte
 * for byte what the 48-byte format delivers; gpet_result_singles_compact returns them as they arrived.  Refused
 * (GPET_ERR_ARG from gpet_run) where the identity would not hold: PSF input (ids come from the file), noise singles, panel
 * ids above 255, more than 4096 modules per panel or 2048 crystals per module; ignored by file runs. */
enum { GPET_SINGLES_RECORDS = 0, GPET_SINGLES_COMPACT = 1 };
typedef struct gpet_single_compact {
    double t;
    float E, x, y, z;
    int32_t eventid;
    uint32_t ids;
} gpet_single_compact;
int gpet_set_singles_format(gpet_ctx* ctx, int format);
int64_t gpet_result_singles_compact(gpet_ctx* ctx, const gpet_single_compact** ptr);
/* The same expansion for records the caller kept (e.g. stored compact): n compact singles -> n 48-byte Events, with the
 * geometry and the digitizer parameters (dead-time level, readout depth / policy) loaded in ctx.  Host only: works on a
 * context without a device. */
int gpet_expand_singles(const gpet_ctx* ctx, const gpet_single_compact* in, int64_t n, gpet_event* out);
/* Phase-space dumps of gpet_run(output_dir) = the reference's OUTPUTPSF switch (constants.h:5; gPET.cu:63-114, 296-351):
 * 0 none; 1 PSF-input mode: the photons entering the phantom -> outsource.dat / idsource.dat / timesource.dat;
 * 2 source mode: those three after source sampling, and in both modes outphantom.dat / idphantom.dat / timephantom.dat
 * after the phantom.  Per live photon (t > 0): 7 x float32 (x y z vx vy vz E), int32 eventid, float64 t -- the layout
 * output/readOutput.m:36-54 reads.  With dumps on, gpet_run materialises the stage queues (staged kernels instead of the
 * fused front end); the photons are the same. */
int gpet_set_psf_output(gpet_ctx* ctx, int mode);
/* *ptr = pairs (2 x uint32 each: earlier single, later single); returns their number (0 in GPET_COINC_RECORDS mode). */
int64_t gpet_result_coincidence_pairs(gpet_ctx* ctx, const uint32_t** ptr);
/* *ptr = one class byte per coincidence of the run, in the order of gpet_result_coincidences / _pairs (0 true,
 * 1 scatter, 2 random; see gpet_fetch_coincidence_classes); returns their number.  gpet_run(output_dir) also appends
 * them to coincidences_class.dat.  The totals are the trues, scatters and randoms fields of gpet_stats. */
int64_t gpet_result_coincidence_classes(gpet_ctx* ctx, const uint8_t** ptr);
int gpet_get_stats(const gpet_ctx* ctx, gpet_stats* stats);
/* Energy spectrum tally of the accumulated singles (nbins over [emin, emax)), kept on device during the run;
 * this is what multi-GPU runs all-reduce. */
int gpet_get_spectrum(gpet_ctx* ctx, uint64_t* bins, int nbins);
int gpet_set_spectrum(gpet_ctx* ctx, int nbins, float emin, float emax);
/* Per-kernel device times (CUDA events bracketing every launch of this library's kernels on the launching stream).
 * Replaces the reference's clock() prints (main.cu:39-41, gPET.cu:245-247, 432-435).  Off by default; enabling clears
 * the accumulated numbers.  gpet_profile_count synchronises the stream and returns the number of distinct kernels. */
int gpet_profile_enable(gpet_ctx* ctx, int on);
int gpet_profile_count(gpet_ctx* ctx);
int gpet_profile_get(gpet_ctx* ctx, int i, char* name, int name_cap, double* total_ms, uint64_t* launches);
/* Shard the planned frames: this context only runs frames f with f % world == rank. */
int gpet_set_shard(gpet_ctx* ctx, int rank, int world);
/* Multi-GPU exchange by time slice (gpet_b200/multi.py).  Decay histories shard over the GPUs; what couples them
 * again is the digitizer -- dead time per site and the coincidence sorter over the global time order (the reference
 * digitizes each epoch = time slice as one list, gPET.cu:385-424).  Every rank therefore receives the post-readout events
 * of ITS time slice [lo, hi) from all ranks plus a halo of the neighbouring slices, digitizes them as one list and keeps
 * what belongs to the slice: the singles with lo <= t < hi (a contiguous range of the time-sorted list: counts[0] singles
 * precede it, counts[1] are inside) and the coincidences whose opening single lies in the slice.  halo_start_us = time
 * from which on the list is complete (-INFINITY: from the start of the acquisition); decisions that would need anything
 * earlier raise counts[2] (halo too short) instead of being silently wrong. */
int gpet_set_emit_window(gpet_ctx* ctx, double lo_us, double hi_us, double halo_start_us);
int gpet_clear_emit_window(gpet_ctx* ctx);
int gpet_get_emit_counts(gpet_ctx* ctx, uint64_t counts[4]);   /* before, inside, halo too short, coincidences emitted */
/* The event buffer <-> caller-owned DEVICE memory (48-byte records), for exchanges between GPUs that never touch the host. */
int64_t gpet_copy_events_to_device(gpet_ctx* ctx, void* dst_device, int64_t cap);
int gpet_put_events_device(gpet_ctx* ctx, const void* src_device, int64_t n);
/* 64-bit history numbers.  The reference indexes atoms and threads with 32-bit integers (gPET.h:50 `unsigned int natom`,
 * gPET_kernals.cu:490-497), which caps an acquisition near 4e9 histories.  Here every pair has a 64-bit global index
 * (pair k of the acquisition = first_pair + k; photons 2k, 2k+1) that keys all its Philox streams; the 32-bit
 * eventid / parn fields of the output records keep the low 31 bits (parn == -1 stays the mark of a noise single).
 * first_pair (default 0) lets a caller continue, or shard by decay index, a longer history sequence. */
int gpet_set_first_pair(gpet_ctx* ctx, uint64_t first_pair);
/* "GPU index" line of input_PET.in (main.cu:52-56) without creating a context: the device a CLI run should use.
 * Returns the index (>= 0) or a negative gpet_status. */
int gpet_peek_config_device(const char* input_file);
/* The direction table gpet_run narrows the panel search with: 32^3 words over the direction cube
 * [-1,1]^3, cell = floor((v + 1) * 16) per axis, x fastest; bit i set = a photon flying in a direction of that cell whose
 * line passes the reference sphere (centre x y z, radius -> ref_sphere) can enter panel i.  Returns the number of words
 * (0: the table does not apply to the loaded inputs -- no phantom/geometry, more than 32 panels, positron range on).
 * Works on a host-only context; exists so that its conservativeness can be tested without a GPU. */
int64_t gpet_get_direction_table(const gpet_ctx* ctx, uint32_t* out, int64_t cap, double ref_sphere[4]);

#ifdef __cplusplus
}
#endif
#endif /* GPET_B200_H */
